"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the
product package (dumux_b200).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ORACLE_LIB: bench.py's CPU timing legs load the -O3 -march=native build (oracle/_fast/liboracle_fast.so, `make fast`) -- the
# reference's own optimisation flags (cmake.opts:17-27), compiler-chosen FMA; parity tests always use the canonical build
_LIB_PATH = os.environ.get("ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle.h", "det_math.h", "Makefile")]
    if os.environ.get("ORACLE_LIB"):
        return _LIB_PATH
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return _LIB_PATH


class OrcOptions(C.Structure):
    _fields_ = [("enable_gravity", C.c_int), ("gravity", C.c_double), ("upwind_weight", C.c_double),
                ("fd_method", C.c_int), ("base_eps", C.c_double), ("privar_magnitude", C.c_double * 2),
                ("stationary", C.c_int), ("dt", C.c_double), ("extrusion", C.c_double),
                ("use_std_pow", C.c_int), ("num_threads", C.c_int)]


class OrcNewtonOptions(C.Structure):
    _fields_ = [("use_line_search", C.c_int), ("line_search_min_relaxation", C.c_double),
                ("enable_shift_criterion", C.c_int), ("enable_residual_criterion", C.c_int),
                ("enable_absolute_residual_criterion", C.c_int), ("satisfy_residual_and_shift", C.c_int),
                ("residual_reduction", C.c_double), ("max_absolute_residual", C.c_double)]


class OrcNewtonReport(C.Structure):
    _fields_ = [("newton_iterations", C.c_int), ("converged", C.c_int), ("linear_iterations_total", C.c_int),
                ("last_shift", C.c_double), ("t_assemble", C.c_double), ("t_solve", C.c_double),
                ("t_update", C.c_double), ("linear_iterations", C.c_int * 64), ("shifts", C.c_double * 64)]


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp]
        L.orc_create_tensor.restype = vp
        L.orc_create_tensor.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, _dp]
        L.orc_destroy.argtypes = [vp]
        L.orc_default_options.argtypes = [C.POINTER(OrcOptions)]
        L.orc_set_options.argtypes = [vp, C.POINTER(OrcOptions)]
        L.orc_set_cell_fields.argtypes = [vp, _dp, _dp, _ip]
        L.orc_set_source.argtypes = [vp, _dp]
        L.orc_set_permeability_diagonal.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_material.argtypes = [vp, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int, _dp]
        L.orc_set_wetting_phase.argtypes = [vp, C.c_int, C.c_int]
        L.orc_set_fluids.argtypes = [vp, _dp, _dp]
        L.orc_set_fluid_table.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, C.c_double]
        L.orc_side_faces.argtypes = [vp, C.c_int]
        L.orc_side_faces.restype = C.c_int
        L.orc_set_boundary.argtypes = [vp, C.c_int, _ip, _dp]
        L.orc_pattern_nnz.argtypes = [vp]
        L.orc_pattern_nnz.restype = C.c_int
        L.orc_pattern.argtypes = [vp, _ip, _ip]
        L.orc_assemble.argtypes = [vp, _dp, vp, vp, vp]
        L.orc_volvars.argtypes = [vp, _dp, _dp]
        L.orc_default_newton_options.argtypes = [C.c_void_p]
        L.orc_newton_solve_ex.argtypes = [vp, _dp, vp, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, _dp, C.c_void_p]
        L.orc_newton_solve_ex.restype = C.c_int
        L.orc_volume_flux.argtypes = [vp, _dp, _dp]
        L.orc_tracer_assemble.argtypes = [vp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_ilu0_bicgstab.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double, C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_ilu0_bicgstab.restype = C.c_int
        L.orc_ilu0_gmres.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_ilu0_gmres.restype = C.c_int
        L.orc_set_linear_solver.argtypes = [vp, C.c_int, C.c_int]
        L.orc_set_tracer_diffusion.argtypes = [vp, C.c_double, C.c_double]
        L.orc_set_tracer_dispersion.argtypes = [vp, C.c_void_p]
        L.orc_ssor_solve.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int, C.c_double, C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_ssor_solve.restype = C.c_int
        L.orc_ssor_apply.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.orc_parmt_solve.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int, C.c_double, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_parmt_solve.restype = C.c_int
        L.orc_parmt_apply.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.orc_parmt_colors.argtypes = [C.c_int, _ip, _ip, _ip]
        L.orc_amg_galerkin.argtypes = [C.c_int, C.c_int, _ip, _ip, _ip, _ip, _dp, _ip, _ip, _ip, _dp]
        L.orc_ssor_factor.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp]
        L.orc_ssor_factor.restype = C.c_int
        L.orc_parmt_colors.restype = C.c_int
        L.orc_ilu0_factor.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp]
        L.orc_ilu0_factor.restype = C.c_int
        L.orc_ilu0_apply.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.orc_spmv.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.orc_norm2.argtypes = [C.c_int, _dp]
        L.orc_norm2.restype = C.c_double
        L.orc_dot.argtypes = [C.c_int, _dp, _dp]
        L.orc_dot.restype = C.c_double
        L.orc_max_relative_shift.argtypes = [C.c_int, _dp, _dp]
        L.orc_max_relative_shift.restype = C.c_double
        L.orc_newton_solve.argtypes = [vp, _dp, vp, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int,
                                       C.POINTER(OrcNewtonReport)]
        L.orc_newton_solve.restype = C.c_int
        L.orc_run_timeloop.argtypes = [vp, _dp, C.c_double, C.c_double, C.c_double, _ip, _dp, C.c_int]
        L.orc_run_timeloop.restype = C.c_int
        L.orc_law_eval.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        L.orc_law_eval.restype = C.c_double
        L.orc_pow.argtypes = [C.c_double, C.c_double]
        L.orc_pow.restype = C.c_double
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU oracle instance for one ProblemSpec (dumux_b200.problems.ProblemSpec duck type)."""

    def __init__(self, spec, use_std_pow: bool = False, num_threads: int = 0):
        L = lib()
        self.spec = spec
        self.b = spec.num_eq
        self.n = spec.num_cells
        cells = np.ascontiguousarray(spec.cells, dtype=np.int32)
        lower = np.ascontiguousarray(spec.lower, dtype=np.float64)
        upper = np.ascontiguousarray(spec.upper, dtype=np.float64)
        nodes = getattr(spec, "node_coords", None)
        if nodes is not None:
            # explicit node coordinates per axis (a slab of a larger YaspGrid keeps the GLOBAL coordinates origin + i*h)
            xs = [np.ascontiguousarray(nodes[a], dtype=np.float64) if a < spec.dim else np.array([0.0, 1.0]) for a in range(3)]
            self.h = C.c_void_p(L.orc_create_tensor(min(spec.model, 2) if spec.model != 3 else 1, spec.dim, cells, xs[0], xs[1], xs[2]))
        else:
            self.h = C.c_void_p(L.orc_create(1 if spec.model == 3 else spec.model, spec.dim, cells, lower, upper))
        self.opt = OrcOptions()
        L.orc_default_options(C.byref(self.opt))
        o = spec.options
        self.opt.enable_gravity = int(o.enable_gravity)
        self.opt.gravity = o.gravity
        self.opt.upwind_weight = o.upwind_weight
        self.opt.fd_method = o.fd_method
        self.opt.base_eps = o.base_eps
        self.opt.privar_magnitude[0] = o.privar_magnitude[0]
        self.opt.privar_magnitude[1] = o.privar_magnitude[1]
        self.opt.stationary = int(o.stationary)
        self.opt.dt = o.dt
        self.opt.extrusion = o.extrusion
        self.opt.use_std_pow = int(use_std_pow)
        self.opt.num_threads = num_threads
        L.orc_set_options(self.h, C.byref(self.opt))
        K = np.asarray(spec.K, dtype=np.float64)
        L.orc_set_cell_fields(self.h, np.ascontiguousarray(K if K.ndim == 1 else K[:, spec.dim - 1]),
                              np.ascontiguousarray(spec.phi, dtype=np.float64),
                              np.ascontiguousarray(spec.region, dtype=np.int32))
        if K.ndim == 2:          # diagonal permeability tensor: K[:, a] = K_aa
            ks = [np.ascontiguousarray(K[:, a]) for a in range(spec.dim)]
            ptr = [k.ctypes.data_as(C.c_void_p) for k in ks] + [None] * (3 - spec.dim)
            L.orc_set_permeability_diagonal(self.h, *ptr)
        for r, m in enumerate(spec.materials):
            reg = np.ascontiguousarray(m.reg if len(m.reg) else [0.01, 0.99, 0.1, 0.9], dtype=np.float64)
            L.orc_set_material(self.h, r, m.law, np.ascontiguousarray(m.params, dtype=np.float64), m.swr, m.snr,
                               int(m.regularize), reg)
            L.orc_set_wetting_phase(self.h, r, int(m.wetting))
        if spec.model == 3:
            L.orc_set_tracer_diffusion(self.h, float(spec.tracer_diffusion[0]), float(spec.tracer_diffusion[1]))
            if getattr(spec, "tracer_dispersion", None) is not None:
                self._disp = np.ascontiguousarray(spec.tracer_dispersion, dtype=np.float64).reshape(-1)
                L.orc_set_tracer_dispersion(self.h, self._disp.ctypes.data_as(C.c_void_p))
        if spec.fluid_table is not None:
            t = spec.fluid_table
            L.orc_set_fluid_table(self.h, t["nT"], t["nP"], t["Tmin"], t["Tmax"],
                                  np.ascontiguousarray(t["pmin"]), np.ascontiguousarray(t["pmax"]),
                                  np.ascontiguousarray(t["rho"]), np.ascontiguousarray(t["mu"]), t["T"])
        else:
            L.orc_set_fluids(self.h, np.ascontiguousarray(spec.rho, dtype=np.float64),
                             np.ascontiguousarray(spec.mu, dtype=np.float64))
        for side, t in spec.bc_type.items():
            assert L.orc_side_faces(self.h, side) == t.shape[0]
            L.orc_set_boundary(self.h, side, np.ascontiguousarray(t, dtype=np.int32),
                               np.ascontiguousarray(spec.bc_values[side], dtype=np.float64))
        if spec.source is not None:
            L.orc_set_source(self.h, np.ascontiguousarray(spec.source, dtype=np.float64))
        self.nnzb = L.orc_pattern_nnz(self.h)
        self.rowptr = np.zeros(self.n + 1, dtype=np.int32)
        self.colidx = np.zeros(self.nnzb, dtype=np.int32)
        L.orc_pattern(self.h, self.rowptr, self.colidx)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_dt(self, dt: float):
        self.opt.dt = dt
        lib().orc_set_options(self.h, C.byref(self.opt))

    def set_threads(self, nt: int):
        self.opt.num_threads = nt
        lib().orc_set_options(self.h, C.byref(self.opt))

    def volume_flux(self, pressure):
        """examples/1ptracer/main.cc:162-199: volume fluxes [n, 2*dim] of a 1p problem from its pressure solution"""
        out = np.zeros((self.n, 2 * self.spec.dim))
        lib().orc_volume_flux(self.h, np.ascontiguousarray(pressure, dtype=np.float64).reshape(-1), out)
        return out

    def assemble(self, cur, prev=None, jacobian=True):
        cur = np.ascontiguousarray(cur, dtype=np.float64).reshape(-1)
        prev_a = None if prev is None else np.ascontiguousarray(prev, dtype=np.float64).reshape(-1)
        res = np.zeros(self.n * self.b)
        jac = np.zeros(self.nnzb * self.b * self.b) if jacobian else None
        if self.spec.model == 3:       # tracer transport on a frozen velocity field
            lib().orc_tracer_assemble(self.h, np.ascontiguousarray(self.spec.volume_flux, dtype=np.float64), cur, prev_a,
                                      int(self.spec.implicit), float(self.spec.rho[0]), _ptr(res), _ptr(jac))
            return res, jac
        lib().orc_assemble(self.h, cur, _ptr(prev_a), _ptr(res), _ptr(jac))
        return res, jac

    def volvars(self, cur):
        cur = np.ascontiguousarray(cur, dtype=np.float64).reshape(-1)
        out = np.zeros((self.n, 12))
        lib().orc_volvars(self.h, cur, out)
        return out

    def set_linear_solver(self, kind, restart=10):
        """kind: 'bicgstab' (ILUBiCGSTABIstlSolver) or 'gmres' (ILURestartedGMResIstlSolver); used by newton()/run_timeloop()."""
        lib().orc_set_linear_solver(self.h, {"bicgstab": 0, "gmres": 1}[kind], restart)

    def solve_gmres(self, values, rhs, reduction=1e-6, maxit=250, restart=10, x0=None):
        x = np.zeros(self.n * self.b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        its, red = C.c_int(0), C.c_double(0)
        st = lib().orc_ilu0_gmres(self.n, self.b, self.rowptr, self.colidx, np.ascontiguousarray(values),
                                  x, np.ascontiguousarray(rhs), reduction, maxit, restart, C.byref(its), C.byref(red))
        return x, st, its.value, red.value

    def solve(self, values, rhs, reduction=1e-6, maxit=250, x0=None):
        x = np.zeros(self.n * self.b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        its, red = C.c_int(0), C.c_double(0)
        st = lib().orc_ilu0_bicgstab(self.n, self.b, self.rowptr, self.colidx, np.ascontiguousarray(values),
                                     x, np.ascontiguousarray(rhs), reduction, maxit, C.byref(its), C.byref(red))
        return x, st, its.value, red.value

    def newton(self, u, prev, lin_reduction=1e-6, lin_maxit=250, max_rel_shift=1e-8, min_steps=2, max_steps=18):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1).copy()
        prev_a = None if prev is None else np.ascontiguousarray(prev, dtype=np.float64).reshape(-1)
        rep = OrcNewtonReport()
        st = lib().orc_newton_solve(self.h, u, _ptr(prev_a), lin_reduction, lin_maxit, max_rel_shift, min_steps,
                                    max_steps, C.byref(rep))
        return u, st, rep

    def newton_ex(self, u, prev, lin_reduction=1e-6, lin_maxit=250, max_rel_shift=1e-8, min_steps=2, max_steps=18, **opts):
        """Newton solve with line search / residual criteria (orc_newton_solve_ex); opts = fields of orc_newton_options.
        Returns (u, status, report, relaxation factors per iteration, last residual reduction)."""
        L = lib()
        o = OrcNewtonOptions()
        L.orc_default_newton_options(C.byref(o))
        for k, v in opts.items():
            setattr(o, k, v)
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1).copy()
        prev_a = None if prev is None else np.ascontiguousarray(prev, dtype=np.float64).reshape(-1)
        rep = OrcNewtonReport()
        relax = np.zeros(64)
        red = C.c_double(0)
        st = L.orc_newton_solve_ex(self.h, u, _ptr(prev_a), lin_reduction, lin_maxit, max_rel_shift, min_steps, max_steps,
                                   C.byref(o), C.byref(rep), relax, C.byref(red))
        return u, st, rep, relax[:rep.newton_iterations].copy(), red.value

    def run_instationary(self, u0, loop, **newton_kw):
        """The same driver as Engine.run_instationary (dumux_b200.timeloop decides dt), Newton steps on the CPU."""
        from dumux_b200 import timeloop
        orc = self
        state = {"u": np.ascontiguousarray(u0, dtype=np.float64).copy(), "prev": np.ascontiguousarray(u0, dtype=np.float64).copy()}

        class _Stepper:
            def solve(self, dt):
                orc.set_dt(dt)
                u, st, rep = orc.newton(state["u"], state["prev"], **newton_kw)
                state["u"] = u
                return st == 0, rep.newton_iterations

            def reset(self):
                state["u"] = state["prev"].copy()

            def advance(self):
                state["prev"] = state["u"].copy()

        its, dts = timeloop.run_instationary(_Stepper(), loop)
        return state["u"], its, dts

    def run_timeloop(self, u0, t_end, dt_initial, max_dt=1e300, max_steps_out=4096):
        u = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1).copy()
        its = np.zeros(max_steps_out, dtype=np.int32)
        dts = np.zeros(max_steps_out)
        nsteps = lib().orc_run_timeloop(self.h, u, t_end, dt_initial, max_dt, its, dts, max_steps_out)
        return u, nsteps, its[:max(nsteps, 0)], dts[:max(nsteps, 0)]

    def law(self, region, which, sw):
        return lib().orc_law_eval(self.h, region, which, float(sw))


def ssor_apply(n, b, rowptr, colidx, values, d):
    v = np.zeros(n * b)
    lib().orc_ssor_apply(n, b, rowptr, colidx, np.ascontiguousarray(values), v, np.ascontiguousarray(d))
    return v


def ssor_solve(n, b, rowptr, colidx, values, rhs, krylov="cg", reduction=1e-6, maxit=250, x0=None):
    """SSORCGIstlSolver ('cg') / SSORBiCGSTABIstlSolver ('bicgstab'); returns (x, status, iterations, reduction)."""
    x = np.zeros(n * b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
    its, red = C.c_int(0), C.c_double(0)
    st = lib().orc_ssor_solve(n, b, rowptr, colidx, np.ascontiguousarray(values), x, np.ascontiguousarray(rhs),
                              {"cg": 0, "bicgstab": 1}[krylov], reduction, maxit, C.byref(its), C.byref(red))
    return x, st, its.value, red.value


PARMT_JAC, PARMT_SOR, PARMT_SSOR = 0, 1, 2


def parmt_apply(kind, n, b, rowptr, colidx, values, d, iterations=1, relaxation=1.0):
    """v = ParMTJac / ParMTSOR / ParMTSSOR (dumux/linear/preconditioners.hh:330-620) applied to d from v = 0"""
    v = np.zeros(n * b)
    lib().orc_parmt_apply(kind, iterations, relaxation, n, b, rowptr, colidx, np.ascontiguousarray(values, dtype=np.float64),
                          v, np.ascontiguousarray(d, dtype=np.float64))
    return v


def parmt_solve(kind, n, b, rowptr, colidx, values, rhs, krylov="bicgstab", reduction=1e-6, maxit=250, iterations=1, relaxation=1.0, x0=None):
    x = np.zeros(n * b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
    its, red = C.c_int(0), C.c_double(0)
    st = lib().orc_parmt_solve(kind, iterations, relaxation, n, b, rowptr, colidx, np.ascontiguousarray(values, dtype=np.float64), x,
                               np.ascontiguousarray(rhs, dtype=np.float64), 0 if krylov == "cg" else 1, reduction, maxit,
                               C.byref(its), C.byref(red))
    return x, st, its.value, red.value


def parmt_colors(n, rowptr, colidx):
    colors = np.zeros(n, dtype=np.int32)
    nc = lib().orc_parmt_colors(n, rowptr, colidx, colors)
    return colors, nc


def ssor_factor(n, b, rowptr, colidx, values):
    """SeqSSOR in factorised form (D + L) D^-1 (D + U), in the layout ilu0_apply consumes; returns (values, status)"""
    out = np.empty(len(values))
    st = lib().orc_ssor_factor(n, b, rowptr, colidx, np.ascontiguousarray(values, dtype=np.float64), out)
    return out, st


def ilu0_factor(n, b, rowptr, colidx, values):
    ilu = np.zeros_like(values)
    st = lib().orc_ilu0_factor(n, b, rowptr, colidx, np.ascontiguousarray(values), ilu)
    return ilu, st


def ilu0_apply(n, b, rowptr, colidx, ilu, d):
    v = np.zeros(n * b)
    lib().orc_ilu0_apply(n, b, rowptr, colidx, np.ascontiguousarray(ilu), v, np.ascontiguousarray(d))
    return v


def spmv(n, b, rowptr, colidx, values, x):
    y = np.zeros(n * b)
    lib().orc_spmv(n, b, rowptr, colidx, np.ascontiguousarray(values), np.ascontiguousarray(x), y)
    return y
