"""ctypes binding of libdumux_b200.so -- the thin host side used by tests and bench.py.

Mirrors the reference's call sequence (FVAssembler::assembleJacobianAndResidual, ILUBiCGSTABIstlSolver::solve,
NewtonSolver::solve) on top of the C ABI declared in include/dumux_b200.h.  There is no CPU fallback: if the
CUDA library is missing or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMX_LIB") or os.path.join(_HERE, "libdumux_b200.so")       # DMX_LIB: developer A/B builds

VEC_CUR, VEC_PREV, VEC_RESIDUAL, VEC_DELTA, VEC_ULAST, VEC_WORK0, VEC_WORK1 = range(7)
PRECOND_ILU0, PRECOND_BLOCKJACOBI = 0, 1
STATUS_OK, STATUS_NOT_CONVERGED, STATUS_BREAKDOWN, STATUS_NONFINITE = 0, 1, 2, 3
KERNEL_ASSEMBLY, KERNEL_SPMV, KERNEL_ILU_APPLY, KERNEL_ILU_FACTOR = range(4)

EXPORTS = [
    "dmx_default_options", "dmx_default_newton_params", "dmx_create", "dmx_create_distributed", "dmx_get_nccl_unique_id",
    "dmx_destroy", "dmx_last_error", "dmx_version", "dmx_grid_structured", "dmx_grid_tensor", "dmx_local_box",
    "dmx_local_box3", "dmx_set_partitioning", "dmx_set_preconditioner_params", "dmx_precond_apply",
    "dmx_default_amg_params", "dmx_set_amg_params", "dmx_amg_level_profile", "dmx_amg_levels", "dmx_amg_level_cells", "dmx_amg_level_nnz_blocks", "dmx_amg_level_matrix",
    "dmx_num_cells", "dmx_num_eq", "dmx_nnz_blocks", "dmx_pattern", "dmx_bcrs_pattern", "dmx_set_options",
    "dmx_set_cell_fields", "dmx_set_permeability_diagonal", "dmx_set_source", "dmx_set_material", "dmx_set_fluids", "dmx_set_fluid_table",
    "dmx_side_faces", "dmx_set_boundary", "dmx_vec_upload", "dmx_vec_download", "dmx_vec_copy", "dmx_jacobian_upload",
    "dmx_jacobian_download", "dmx_vec_device_ptr", "dmx_jacobian_device_ptr", "dmx_assemble", "dmx_assemble_host",
    "dmx_linear_solve", "dmx_linear_solve_host", "dmx_norm2", "dmx_newton_update", "dmx_newton_solve",
    "dmx_newton_solve_host", "dmx_newton_step", "dmx_advance_timestep", "dmx_reset_timestep", "dmx_spmv",
    "dmx_ilu0_factor", "dmx_ilu0_apply", "dmx_ilu0_download", "dmx_dot", "dmx_halo_exchange", "dmx_time_kernel",
    "dmx_kernel_launch_count", "dmx_synchronize", "dmx_profile", "dmx_profile_read",
    "dmx_newton_step_host", "dmx_timer_start", "dmx_timer_stop", "dmx_debug_sweep_trace",
    "dmx_volume_flux", "dmx_set_volume_flux", "dmx_set_tracer", "dmx_set_wetting_phase", "dmx_set_linear_solver", "dmx_ssor_apply", "dmx_num_output_fields", "dmx_output_fields", "dmx_set_tracer_diffusion", "dmx_set_tracer_dispersion",
]
SOLVER_BICGSTAB, SOLVER_RESTARTED_GMRES, SOLVER_CG = 0, 1, 2
PRECOND_SSOR = 2
PRECOND_PARMT_JAC, PRECOND_PARMT_SOR, PRECOND_PARMT_SSOR = 3, 4, 5
PRECOND_AMG = 6
K_ASSEMBLY, K_SPMV, K_ILU_APPLY, K_ILU_FACTOR, K_AMG, K_BLAS1, K_HALO, K_JACOBI = range(8)


class DmxOptions(C.Structure):
    _fields_ = [("enable_gravity", C.c_int), ("gravity", C.c_double), ("upwind_weight", C.c_double),
                ("fd_method", C.c_int), ("base_eps", C.c_double), ("privar_magnitude", C.c_double * 2),
                ("stationary", C.c_int), ("dt", C.c_double), ("extrusion", C.c_double)]


class DmxNewtonParams(C.Structure):
    _fields_ = [("max_relative_shift", C.c_double), ("min_steps", C.c_int), ("max_steps", C.c_int),
                ("lin_reduction", C.c_double), ("lin_maxit", C.c_int), ("preconditioner", C.c_int),
                ("use_line_search", C.c_int), ("line_search_min_relaxation", C.c_double),
                ("enable_shift_criterion", C.c_int), ("enable_residual_criterion", C.c_int),
                ("enable_absolute_residual_criterion", C.c_int), ("satisfy_residual_and_shift", C.c_int),
                ("residual_reduction", C.c_double), ("max_absolute_residual", C.c_double)]


class DmxAmgParams(C.Structure):
    _fields_ = [("pre_steps", C.c_int), ("post_steps", C.c_int), ("prolongation_damping", C.c_double), ("smoother", C.c_int),
                ("coarsest_cells", C.c_int), ("coarsest_steps", C.c_int), ("max_levels", C.c_int),
                ("smoother_iterations", C.c_int), ("smoother_relaxation", C.c_double)]


class DmxNewtonReport(C.Structure):
    _fields_ = [("newton_iterations", C.c_int), ("converged", C.c_int), ("linear_iterations_total", C.c_int),
                ("last_shift", C.c_double), ("t_assemble", C.c_double), ("t_solve", C.c_double),
                ("t_update", C.c_double), ("linear_iterations", C.c_int * 64), ("shifts", C.c_double * 64),
                ("last_reduction", C.c_double), ("last_residual_norm", C.c_double), ("relaxation", C.c_double * 64)]


class DmxError(RuntimeError):
    pass


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def load_library():
    """Load libdumux_b200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DmxError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.dmx_last_error.restype = C.c_char_p
    L.dmx_last_error.argtypes = [vp]
    L.dmx_version.restype = C.c_char_p
    L.dmx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.dmx_create_distributed.argtypes = [C.POINTER(vp), C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.dmx_get_nccl_unique_id.argtypes = [C.c_void_p]
    L.dmx_destroy.argtypes = [vp]
    L.dmx_default_options.argtypes = [C.POINTER(DmxOptions)]
    L.dmx_default_newton_params.argtypes = [C.POINTER(DmxNewtonParams)]
    L.dmx_grid_structured.argtypes = [vp, C.c_int, C.c_int, _ip, _dp, _dp]
    L.dmx_local_box.argtypes = [vp, _ip, _ip, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.dmx_local_box3.argtypes = [vp, _ip, _ip, _ip, _ip, _ip, _ip]
    L.dmx_set_partitioning.argtypes = [vp, C.c_void_p]
    L.dmx_num_cells.argtypes = [vp]
    L.dmx_num_eq.argtypes = [vp]
    L.dmx_nnz_blocks.argtypes = [vp]
    L.dmx_nnz_blocks.restype = C.c_longlong
    L.dmx_pattern.argtypes = [vp, _ip, _ip]
    L.dmx_bcrs_pattern.argtypes = [vp, C.c_int, C.c_int, _ip, _ip]
    L.dmx_set_options.argtypes = [vp, C.POINTER(DmxOptions)]
    L.dmx_set_cell_fields.argtypes = [vp, _dp, _dp, _ip]
    L.dmx_set_permeability_diagonal.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dmx_set_source.argtypes = [vp, _dp]
    L.dmx_set_material.argtypes = [vp, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int, _dp]
    L.dmx_set_fluids.argtypes = [vp, _dp, _dp]
    L.dmx_set_fluid_table.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, C.c_double]
    L.dmx_side_faces.argtypes = [vp, C.c_int]
    L.dmx_volume_flux.argtypes = [vp, _dp]
    L.dmx_set_volume_flux.argtypes = [vp, _dp]
    L.dmx_set_tracer.argtypes = [vp, C.c_int]
    L.dmx_set_tracer_diffusion.argtypes = [vp, C.c_double, C.c_double]
    L.dmx_set_tracer_dispersion.argtypes = [vp, C.c_void_p]
    L.dmx_set_wetting_phase.argtypes = [vp, C.c_int, C.c_int]
    L.dmx_set_linear_solver.argtypes = [vp, C.c_int, C.c_int]
    L.dmx_ssor_apply.argtypes = [vp, C.c_int, C.c_int]
    L.dmx_set_preconditioner_params.argtypes = [vp, C.c_int, C.c_double]
    L.dmx_default_amg_params.argtypes = [C.POINTER(DmxAmgParams)]
    L.dmx_set_amg_params.argtypes = [vp, C.POINTER(DmxAmgParams)]
    L.dmx_amg_levels.argtypes = [vp]
    L.dmx_amg_level_cells.argtypes = [vp, C.c_int, _ip]
    L.dmx_amg_level_nnz_blocks.argtypes = [vp, C.c_int]
    L.dmx_amg_level_nnz_blocks.restype = C.c_longlong
    L.dmx_amg_level_matrix.argtypes = [vp, C.c_int, C.c_void_p]
    L.dmx_precond_apply.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.dmx_num_output_fields.argtypes = [vp]
    L.dmx_output_fields.argtypes = [vp, C.c_void_p]
    L.dmx_set_boundary.argtypes = [vp, C.c_int, _ip, _dp]
    L.dmx_vec_upload.argtypes = [vp, C.c_int, C.c_void_p]
    L.dmx_vec_download.argtypes = [vp, C.c_int, C.c_void_p]
    L.dmx_vec_copy.argtypes = [vp, C.c_int, C.c_int]
    L.dmx_jacobian_upload.argtypes = [vp, C.c_void_p]
    L.dmx_jacobian_download.argtypes = [vp, C.c_void_p]
    L.dmx_vec_device_ptr.argtypes = [vp, C.c_int]
    L.dmx_vec_device_ptr.restype = C.c_void_p
    L.dmx_jacobian_device_ptr.argtypes = [vp]
    L.dmx_jacobian_device_ptr.restype = C.c_void_p
    L.dmx_assemble.argtypes = [vp, C.c_int]
    L.dmx_assemble_host.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dmx_linear_solve.argtypes = [vp, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.dmx_linear_solve_host.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                        C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.dmx_norm2.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.dmx_newton_update.argtypes = [vp, C.POINTER(C.c_double)]
    L.dmx_newton_solve.argtypes = [vp, C.POINTER(DmxNewtonParams), C.POINTER(DmxNewtonReport)]
    L.dmx_newton_solve_host.argtypes = [vp, C.c_void_p, C.c_void_p, C.POINTER(DmxNewtonParams), C.POINTER(DmxNewtonReport)]
    L.dmx_newton_step.argtypes = [vp, C.POINTER(DmxNewtonParams), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                  C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.dmx_advance_timestep.argtypes = [vp]
    L.dmx_reset_timestep.argtypes = [vp]
    L.dmx_spmv.argtypes = [vp, C.c_int, C.c_int]
    L.dmx_ilu0_factor.argtypes = [vp]
    L.dmx_ilu0_apply.argtypes = [vp, C.c_int, C.c_int]
    L.dmx_ilu0_download.argtypes = [vp, C.c_void_p]
    L.dmx_dot.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.dmx_halo_exchange.argtypes = [vp, C.c_int]
    L.dmx_time_kernel.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.dmx_kernel_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.dmx_synchronize.argtypes = [vp]
    L.dmx_newton_step_host.argtypes = [vp, C.c_void_p, C.POINTER(DmxNewtonParams), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.dmx_debug_sweep_trace.argtypes = [vp, C.c_void_p]
    L.dmx_timer_start.argtypes = [vp]
    L.dmx_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.dmx_profile.argtypes = [vp, C.c_int]
    L.dmx_profile_read.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.dmx_amg_level_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    _lib = L
    return L


def _hostptr(a):
    """Pointer of a numpy array or a (pinned) torch CPU tensor, or None."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One dmx_ctx (one GPU).  `spec` is a dumux_b200.problems.ProblemSpec describing the GLOBAL problem; in a
    distributed engine (`nccl_uid`, `rank`, `nranks`) the per-rank blocks are cut out here.  `part` = Grid.Partitioning
    (ranks per axis, product = nranks; default: slabs along the last axis)."""

    def __init__(self, spec=None, device: int = 0, nccl_uid: bytes | None = None, rank: int = 0, nranks: int = 1, part=None):
        self.L = load_library()
        self.h = C.c_void_p()
        if nranks > 1:
            buf = C.create_string_buffer(nccl_uid, 128)
            rc = self.L.dmx_create_distributed(C.byref(self.h), device, buf, rank, nranks)
        else:
            rc = self.L.dmx_create(C.byref(self.h), device)
        if rc != 0 or not self.h:
            raise DmxError(f"dmx_create failed with code {rc} (is a CUDA device visible? there is no CPU fallback)")
        self.rank, self.nranks = rank, nranks
        if part is not None:
            p3 = np.ascontiguousarray(list(part) + [1] * (3 - len(part)), dtype=np.int32)
            self._check(self.L.dmx_set_partitioning(self.h, p3.ctypes.data_as(C.c_void_p)))
        self.spec = None
        self.n = self.b = 0
        self.opt = DmxOptions()
        self.L.dmx_default_options(C.byref(self.opt))
        if spec is not None:
            self.set_problem(spec)

    # ---- plumbing ----
    def _check(self, rc, allow_status=False):
        if rc < 0 or (rc > 0 and not allow_status):
            raise DmxError(f"libdumux_b200 error {rc}: {self.L.dmx_last_error(self.h).decode()}")
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.L.dmx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def nccl_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        if L.dmx_get_nccl_unique_id(buf) != 0:
            raise DmxError("ncclGetUniqueId failed")
        return buf.raw

    # ---- problem set-up ----
    def set_problem(self, spec):
        L = self.L
        self.spec = spec
        cells = np.ascontiguousarray(spec.cells, dtype=np.int32)
        self._check(L.dmx_grid_structured(self.h, spec.model, spec.dim, cells,
                                          np.ascontiguousarray(spec.lower, dtype=np.float64),
                                          np.ascontiguousarray(spec.upper, dtype=np.float64)))
        self.n = L.dmx_num_cells(self.h)
        self.b = L.dmx_num_eq(self.h)
        self.nnzb = L.dmx_nnz_blocks(self.h)
        lc, off, ob, oe, pp, pc = (np.zeros(3, dtype=np.int32) for _ in range(6))
        L.dmx_local_box3(self.h, lc, off, ob, oe, pp, pc)
        self.local_cells, self.offset, self.own_lo, self.own_hi, self.part, self.coord = lc, off, ob, oe, pp, pc
        self.own_begin, self.own_end = int(ob[spec.dim - 1]), int(oe[spec.dim - 1])      # slab view: the last axis
        self.local_box = [(int(off[a]), int(off[a] + lc[a])) for a in range(spec.dim)]
        sb = spec.local_box
        if sb is not None and [tuple(x) for x in sb] != self.local_box:
            raise DmxError(f"box-local spec covers {sb}, this rank holds {self.local_box}")
        o = spec.options
        self.opt.enable_gravity = int(o.enable_gravity)
        self.opt.gravity = o.gravity
        self.opt.upwind_weight = o.upwind_weight
        self.opt.fd_method = o.fd_method
        self.opt.base_eps = o.base_eps
        self.opt.privar_magnitude[0], self.opt.privar_magnitude[1] = o.privar_magnitude
        self.opt.stationary = int(o.stationary)
        self.opt.dt = o.dt
        self.opt.extrusion = o.extrusion
        self._check(L.dmx_set_options(self.h, C.byref(self.opt)))
        K = self.localize_cells(np.asarray(spec.K, dtype=np.float64))
        self._check(L.dmx_set_cell_fields(self.h, np.ascontiguousarray(K if K.ndim == 1 else K[:, spec.dim - 1]), self.localize_cells(spec.phi),
                                          self.localize_cells(spec.region.astype(np.int32))))
        if K.ndim == 2:          # diagonal permeability tensor: K[:, a] = K_aa (dmx_set_permeability_diagonal)
            ks = [np.ascontiguousarray(K[:, a]) for a in range(spec.dim)]
            ptr = [k.ctypes.data_as(C.c_void_p) for k in ks] + [None] * (3 - spec.dim)
            self._check(L.dmx_set_permeability_diagonal(self.h, *ptr))
        for r, m in enumerate(spec.materials):
            reg = np.ascontiguousarray(m.reg if len(m.reg) else [0.01, 0.99, 0.1, 0.9], dtype=np.float64)
            self._check(L.dmx_set_material(self.h, r, m.law, np.ascontiguousarray(m.params, dtype=np.float64),
                                           m.swr, m.snr, int(m.regularize), reg))
            if getattr(m, "wetting", 0):
                self._check(L.dmx_set_wetting_phase(self.h, r, int(m.wetting)))
        if spec.fluid_table is not None:
            t = spec.fluid_table
            self._check(L.dmx_set_fluid_table(self.h, t["nT"], t["nP"], t["Tmin"], t["Tmax"],
                                              np.ascontiguousarray(t["pmin"]), np.ascontiguousarray(t["pmax"]),
                                              np.ascontiguousarray(t["rho"]), np.ascontiguousarray(t["mu"]), t["T"]))
        else:
            self._check(L.dmx_set_fluids(self.h, np.ascontiguousarray(spec.rho, dtype=np.float64),
                                         np.ascontiguousarray(spec.mu, dtype=np.float64)))
        for side, t in spec.bc_type.items():
            tl, vl = self.localize_side(side, t, spec.bc_values[side])
            self._check(L.dmx_set_boundary(self.h, side, tl, vl))
        if spec.source is not None:
            self._check(L.dmx_set_source(self.h, self.localize_cells(spec.source)))
        if getattr(spec, "volume_flux", None) is not None:
            # slab-decomposed runs: per-cell fluxes of the local box incl. overlap, like every other per-cell array
            self._check(L.dmx_set_volume_flux(self.h, np.ascontiguousarray(self.localize_cells(spec.volume_flux), dtype=np.float64).reshape(-1)))
            self._check(L.dmx_set_tracer(self.h, int(spec.implicit)))
            self._check(L.dmx_set_tracer_diffusion(self.h, float(spec.tracer_diffusion[0]), float(spec.tracer_diffusion[1])))
            if getattr(spec, "tracer_dispersion", None) is not None:
                disp = np.ascontiguousarray(self.localize_cells(spec.tracer_dispersion), dtype=np.float64).reshape(-1)
                self._check(L.dmx_set_tracer_dispersion(self.h, disp.ctypes.data_as(C.c_void_p)))

    def owner_mask(self):
        """bool[n]: cells this rank owns (interior partition), x fastest"""
        m = np.ones((int(self.local_cells[2]), int(self.local_cells[1]), int(self.local_cells[0])), dtype=bool)
        for a in range(3):
            idx = np.arange(int(self.local_cells[a]))
            ok = (idx >= self.own_lo[a]) & (idx < self.own_hi[a])
            shp = [1, 1, 1]
            shp[2 - a] = -1
            m &= ok.reshape(shp)
        return m.reshape(-1)

    def localize_cells(self, a):
        """Cut the local box (incl. overlap) out of a global per-cell array (x fastest)."""
        a = np.asarray(a)
        if self.nranks == 1 or self.spec.local_box is not None:
            return np.ascontiguousarray(a)
        gc = self.spec.cells3
        g = a.reshape((gc[2], gc[1], gc[0]) + a.shape[1:])
        sl = tuple(slice(int(self.offset[d]), int(self.offset[d] + self.local_cells[d])) for d in (2, 1, 0))
        return np.ascontiguousarray(g[sl].reshape((-1,) + a.shape[1:]))

    def localize_side(self, side, t, v):
        t = np.asarray(t, dtype=np.int32)
        v = np.asarray(v, dtype=np.float64)
        if self.nranks == 1 or self.spec.local_box is not None:
            return np.ascontiguousarray(t), np.ascontiguousarray(v)
        gc = self.spec.cells3
        a = side // 2
        # side plane axes: remaining axes ascending, lower fastest (ignored by the library on processor boundaries)
        rem = [d for d in range(3) if d != a]
        shape = (gc[rem[1]], gc[rem[0]])
        sl = tuple(slice(int(self.offset[d]), int(self.offset[d] + self.local_cells[d])) for d in (rem[1], rem[0]))
        tt = t.reshape(shape)[sl]
        vv = v.reshape(shape + (v.shape[-1],))[sl]
        return np.ascontiguousarray(tt.reshape(-1)), np.ascontiguousarray(vv.reshape(-1, v.shape[-1]))

    def set_bcrs_pattern(self, n, b, rowptr, colidx):
        self._check(self.L.dmx_bcrs_pattern(self.h, n, b, np.ascontiguousarray(rowptr, dtype=np.int32),
                                            np.ascontiguousarray(colidx, dtype=np.int32)))
        self.n, self.b, self.nnzb = n, b, int(rowptr[n])

    def volume_flux(self, pressure):
        """examples/1ptracer/main.cc:162-199 on the device: volume fluxes [n, 2*dim] of this 1p problem from `pressure`"""
        self.upload(VEC_CUR, pressure)
        out = np.zeros((self.n, 2 * self.spec.dim))
        self._check(self.L.dmx_volume_flux(self.h, out.reshape(-1)))
        return out

    def set_dt(self, dt):
        self.opt.dt = dt
        self._check(self.L.dmx_set_options(self.h, C.byref(self.opt)))

    def pattern(self):
        rowptr = np.zeros(self.n + 1, dtype=np.int32)
        colidx = np.zeros(self.nnzb, dtype=np.int32)
        self._check(self.L.dmx_pattern(self.h, rowptr, colidx))
        return rowptr, colidx

    # ---- data movement ----
    def upload(self, vec, host):
        if not hasattr(host, "data_ptr"):
            host = np.ascontiguousarray(host, dtype=np.float64).reshape(-1)
            assert host.size == self.n * self.b
        self._check(self.L.dmx_vec_upload(self.h, vec, _hostptr(host)))

    def download(self, vec, out=None):
        if out is None:
            out = np.empty(self.n * self.b)
        self._check(self.L.dmx_vec_download(self.h, vec, _hostptr(out)))
        return out

    def jacobian(self):
        out = np.empty(self.nnzb * self.b * self.b)
        self._check(self.L.dmx_jacobian_download(self.h, _hostptr(out)))
        return out

    def upload_jacobian(self, values):
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        self._check(self.L.dmx_jacobian_upload(self.h, _hostptr(values)))

    # ---- the hot path ----
    def assemble(self, cur, prev=None, jacobian=True):
        """FVAssembler::assembleJacobianAndResidual(curSol) with host buffers; returns (residual, jacobian values)."""
        self.upload(VEC_CUR, cur)
        if prev is not None:
            self.upload(VEC_PREV, prev)
        self._check(self.L.dmx_assemble(self.h, int(jacobian)))
        res = self.download(VEC_RESIDUAL)
        return res, (self.jacobian() if jacobian else None)

    def assemble_device(self, jacobian=True):
        return self._check(self.L.dmx_assemble(self.h, int(jacobian)), allow_status=True)

    def solve(self, values, rhs, reduction=1e-6, maxit=250, precond=PRECOND_ILU0, x0=None):
        """ILUBiCGSTABIstlSolver::solve(A, x, b) with host buffers; returns (x, status, iterations, reduction)."""
        x = np.zeros(self.n * self.b) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
        values = np.ascontiguousarray(values, dtype=np.float64)
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        its, red = C.c_int(0), C.c_double(0)
        st = self._check(self.L.dmx_linear_solve_host(self.h, _hostptr(values), _hostptr(x), _hostptr(rhs), reduction,
                                                      maxit, precond, C.byref(its), C.byref(red)), allow_status=True)
        return x, st, its.value, red.value

    def set_linear_solver(self, kind, restart=10):
        """'bicgstab' = ILUBiCGSTABIstlSolver (default), 'gmres' = ILURestartedGMResIstlSolver (LinearSolver.GMResRestart)."""
        self._check(self.L.dmx_set_linear_solver(self.h, {"bicgstab": SOLVER_BICGSTAB, "gmres": SOLVER_RESTARTED_GMRES, "cg": SOLVER_CG}[kind], restart))

    def output_fields(self, phase_names=("aq", "napl")):
        """Output fields of the state in CUR with the reference's names (IOName::saturation/pressure/density/mobility<FluidSystem>,
        capillaryPressure, porosity; dumux/io/name.hh): an ordered dict name -> array[n]."""
        nf = self.L.dmx_num_output_fields(self.h)
        out = np.zeros((nf, self.n_local if hasattr(self, "n_local") else self.n))
        self._check(self.L.dmx_output_fields(self.h, _hostptr(out)))
        if nf == 1:
            return {"p": out[0]}
        names = []
        for ph in phase_names:
            names += [f"S_{ph}", f"p_{ph}", f"rho_{ph}", f"mob_{ph}"]
        names += ["pc", "porosity"]
        return dict(zip(names, out))

    def ssor_apply(self, d_vec, v_vec):
        self._check(self.L.dmx_ssor_apply(self.h, d_vec, v_vec))

    def set_preconditioner_params(self, iterations=1, relaxation=1.0):
        self._check(self.L.dmx_set_preconditioner_params(self.h, iterations, relaxation))

    def set_amg_params(self, **kw):
        """dune-istl's AMG parameter names: pre_steps, post_steps, prolongation_damping, smoother (PRECOND_SSOR | PRECOND_ILU0 |
        PRECOND_PARMT_*), smoother_iterations, smoother_relaxation, coarsest_cells, coarsest_steps, max_levels; unspecified ones keep
        dune's defaults"""
        p = DmxAmgParams()
        self.L.dmx_default_amg_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        self._check(self.L.dmx_set_amg_params(self.h, C.byref(p)))

    def amg_level_profile(self):
        """[(ms per cycle, cycles)] per level, exclusive of the coarser levels (timed while profile(True))"""
        out = []
        for l in range(self.L.dmx_amg_levels(self.h)):
            ms, n = C.c_double(0.0), C.c_longlong(0)
            self._check(self.L.dmx_amg_level_profile(self.h, l, C.byref(ms), C.byref(n)))
            out.append((ms.value / max(1, n.value), int(n.value)))
        return out

    def amg_levels(self):
        out = []
        for l in range(self.L.dmx_amg_levels(self.h)):
            c = np.zeros(3, dtype=np.int32)
            self._check(self.L.dmx_amg_level_cells(self.h, l, c))
            out.append(tuple(int(x) for x in c))
        return out

    def amg_level_matrix(self, level):
        nnzb = self.L.dmx_amg_level_nnz_blocks(self.h, level)
        out = np.empty(nnzb * self.b * self.b)
        self._check(self.L.dmx_amg_level_matrix(self.h, level, _hostptr(out)))
        return out

    def precond_apply(self, precond, d_vec=VEC_WORK0, v_vec=VEC_WORK1):
        self._check(self.L.dmx_precond_apply(self.h, precond, d_vec, v_vec))

    def solve_device(self, reduction=1e-6, maxit=250, precond=PRECOND_ILU0):
        its, red = C.c_int(0), C.c_double(0)
        st = self._check(self.L.dmx_linear_solve(self.h, reduction, maxit, precond, C.byref(its), C.byref(red)),
                         allow_status=True)
        return st, its.value, red.value

    def norm(self, vec):
        out = C.c_double(0)
        self._check(self.L.dmx_norm2(self.h, vec, C.byref(out)))
        return out.value

    def dot(self, a, b):
        out = C.c_double(0)
        self._check(self.L.dmx_dot(self.h, a, b, C.byref(out)))
        return out.value

    def newton_params(self, **kw):
        p = DmxNewtonParams()
        self.L.dmx_default_newton_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def newton(self, u, prev, **kw):
        """NewtonSolver::solve at fixed dt with host buffers; returns (u, status, report)."""
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1).copy()
        prev_a = None if prev is None else np.ascontiguousarray(prev, dtype=np.float64).reshape(-1)
        p = self.newton_params(**kw)
        rep = DmxNewtonReport()
        st = self._check(self.L.dmx_newton_solve_host(self.h, _hostptr(u), _hostptr(prev_a), C.byref(p), C.byref(rep)),
                         allow_status=True)
        return u, st, rep

    def newton_device(self, **kw):
        p = self.newton_params(**kw)
        rep = DmxNewtonReport()
        st = self._check(self.L.dmx_newton_solve(self.h, C.byref(p), C.byref(rep)), allow_status=True)
        return st, rep

    def newton_step(self, params):
        its, shift = C.c_int(0), C.c_double(0)
        a, s, u = C.c_float(0), C.c_float(0), C.c_float(0)
        st = self._check(self.L.dmx_newton_step(self.h, C.byref(params), C.byref(its), C.byref(shift), C.byref(a),
                                                C.byref(s), C.byref(u)), allow_status=True)
        return st, its.value, shift.value, a.value, s.value, u.value

    def newton_step_host(self, u, params):
        """One Newton iteration with a HOST solution vector (numpy or pinned torch tensor), updated in place."""
        its, shift = C.c_int(0), C.c_double(0)
        st = self._check(self.L.dmx_newton_step_host(self.h, _hostptr(u), C.byref(params), C.byref(its), C.byref(shift)),
                         allow_status=True)
        return st, its.value, shift.value

    def timer_start(self):
        self._check(self.L.dmx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self.L.dmx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def newton_update(self):
        out = C.c_double(0)
        self._check(self.L.dmx_newton_update(self.h, C.byref(out)))
        return out.value

    def advance_timestep(self):
        self._check(self.L.dmx_advance_timestep(self.h))

    def reset_timestep(self):
        self._check(self.L.dmx_reset_timestep(self.h))

    def copy(self, dst, src):
        self._check(self.L.dmx_vec_copy(self.h, dst, src))

    # ---- kernel-level ----
    def spmv(self, x_vec=VEC_WORK0, y_vec=VEC_WORK1):
        self._check(self.L.dmx_spmv(self.h, x_vec, y_vec))

    def ilu0_factor(self):
        return self._check(self.L.dmx_ilu0_factor(self.h), allow_status=True)

    def ilu0_apply(self, d_vec=VEC_WORK0, v_vec=VEC_WORK1):
        self._check(self.L.dmx_ilu0_apply(self.h, d_vec, v_vec))

    def ilu0_values(self):
        out = np.empty(self.nnzb * self.b * self.b)
        self._check(self.L.dmx_ilu0_download(self.h, _hostptr(out)))
        return out

    def halo_exchange(self, vec):
        self._check(self.L.dmx_halo_exchange(self.h, vec))

    def time_kernel(self, which, reps=10):
        ms = C.c_float(0)
        self._check(self.L.dmx_time_kernel(self.h, which, reps, C.byref(ms)))
        return ms.value

    def launches(self):
        out = C.c_longlong(0)
        self.L.dmx_kernel_launch_count(self.h, C.byref(out))
        return out.value

    def profile(self, enable=True):
        self._check(self.L.dmx_profile(self.h, int(enable)))

    def profile_read(self, kclass):
        """(device ms accumulated, timed units) of one K_* kernel class since profile(True)."""
        ms, n = C.c_double(0), C.c_longlong(0)
        self._check(self.L.dmx_profile_read(self.h, kclass, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def sweep_trace(self):
        out = np.zeros((2, 2, 64, 24), dtype=np.int64)
        self._check(self.L.dmx_debug_sweep_trace(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def synchronize(self):
        self._check(self.L.dmx_synchronize(self.h))

    def run_instationary(self, u0, loop, **newton_kw):
        """Instationary run driven by a dumux_b200.timeloop.TimeLoop / CheckPointTimeLoop (e.g. the periodic check points of
        test/porousmediumflow/1p/compressible/instationary/main.cc:118-150).  State stays on the device between steps."""
        from . import timeloop
        eng = self

        class _Stepper:
            def solve(self, dt):
                eng.set_dt(dt)
                st, rep = eng.newton_device(**newton_kw)
                return st == 0, rep.newton_iterations

            def reset(self):
                eng.reset_timestep()

            def advance(self):
                eng.advance_timestep()

        self.upload(VEC_CUR, u0)
        self.upload(VEC_PREV, u0)
        its, dts = timeloop.run_instationary(_Stepper(), loop)
        return self.download(VEC_CUR), its, dts

    def run_timeloop(self, u0, t_end, dt_initial, max_dt=1e300, **newton_kw):
        """Instationary run as in test/porousmediumflow/2p/incompressible/main.cc:126-163: plain TimeLoop
        (dumux/common/timeloop.hh:239-252,320-332,385-411), Newton dt-halving retry (newtonsolver.hh:309-355) and
        suggestTimeStepSize (:784-798).  State stays on the device between steps."""
        self.upload(VEC_CUR, u0)
        self.upload(VEC_PREV, u0)
        time, t_start, base_eps = 0.0, 0.0, 1e-10
        finished = lambda: (t_end - time) < base_eps * (time - t_start)
        max_step = lambda: 0.0 if finished() else min(max_dt, max(0.0, t_end - time))
        dt = min(dt_initial, max_step())
        its, dts = [], []
        target = 10
        while True:
            ok = False
            n_steps = 0
            for i in range(11):
                self.set_dt(dt)
                st, rep = self.newton_device(**newton_kw)
                n_steps = rep.newton_iterations
                if st == 0:
                    ok = True
                    break
                if i < 10:
                    self.reset_timestep()
                    dt = min(dt * 0.5, max_step())
            if not ok:
                raise DmxError("Newton solver didn't converge after 10 time-step divisions")
            self.advance_timestep()
            its.append(n_steps)
            dts.append(dt)
            time += dt
            dt = min(dt, max_step())
            if n_steps > target:
                sugg = dt / (1.0 + (n_steps - target) / target)
            else:
                sugg = dt * (1.0 + (target - n_steps) / target / 1.2)
            dt = min(sugg, max_step())
            if finished():
                break
        return self.download(VEC_CUR), its, dts
