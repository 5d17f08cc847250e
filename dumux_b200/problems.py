"""Problem set-ups (the inputs a DuMux Problem/SpatialParams pair provides), sampled into flat arrays.

In DuMux the user supplies C++ callables (``boundaryTypesAtPos``, ``dirichletAtPos``, ``neumannAtPos``,
``permeability``, ``porosityAtPos``, ``fluidMatrixInteraction`` ...; dumux/common/fvproblem.hh:126-283,
dumux/porousmediumflow/fvspatialparams.hh:83-99).  The B200 path evaluates them ONCE on the host into the
flat per-cell / per-boundary-face arrays below; these arrays are what crosses the C ABI.

Each factory cites the reference test it reproduces.  Cells are numbered x-fastest (YaspGrid), boundary
faces of a side are numbered with the lower remaining axis fastest.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Tuple

import numpy as np

MODEL_1P, MODEL_2P, MODEL_TRACER = 1, 2, 3
LAW_BC, LAW_VG = 0, 1
BC_NEUMANN, BC_DIRICHLET, BC_NONE, BC_OUTFLOW = 0, 1, 2, 3


DIFF_ANALYTIC = 100     # Options.fd_method value for DiffMethod::analytic (DMX_DIFF_ANALYTIC / ORC_DIFF_ANALYTIC)


@dataclasses.dataclass
class Material:
    law: int
    params: Tuple[float, ...]          # BC: (pcEntry, lambda); VG: (alpha, n, l)
    swr: float = 0.0
    snr: float = 0.0
    regularize: bool = True
    reg: Tuple[float, ...] = ()        # BC: (pcLowSwe,), VG: (pcLowSwe, pcHighSwe, krnLowSwe, krwHighSwe)
    wetting: int = 0                   # wetting phase index (spatialParams.wettingPhase): 0 = phase 0 (water), 1 = phase 1


@dataclasses.dataclass
class Options:
    enable_gravity: bool = True
    gravity: float = 9.81
    upwind_weight: float = 1.0
    fd_method: int = 1
    base_eps: float = 1e-10
    privar_magnitude: Tuple[float, float] = (-1.0, -1.0)
    stationary: bool = False
    dt: float = 1.0
    extrusion: float = 1.0


@dataclasses.dataclass
class ProblemSpec:
    name: str
    model: int
    dim: int
    cells: Tuple[int, ...]
    lower: Tuple[float, ...]
    upper: Tuple[float, ...]
    K: np.ndarray
    phi: np.ndarray
    region: np.ndarray
    materials: List[Material]
    rho: Tuple[float, ...]
    mu: Tuple[float, ...]
    bc_type: Dict[int, np.ndarray]      # side -> int32[nf]
    bc_values: Dict[int, np.ndarray]    # side -> float64[nf, numEq]
    options: Options
    initial: np.ndarray                 # float64[n, numEq]
    source: Optional[np.ndarray] = None
    fluid_table: Optional[dict] = None  # tabulated liquid (1p compressible)
    # tracer transport (MODEL_TRACER): frozen volume fluxes [n, 2*dim] (sides -x,+x,-y,+y,-z,+z seen from the cell) and the
    # time discretisation (False: explicit Euler as in examples/1ptracer/main.cc:236, True: implicit)
    volume_flux: Optional[np.ndarray] = None
    implicit: bool = False
    # tracer: (binary diffusion coefficient D, SpatialParams.Tortuosity) of Fick's law with DiffusivityConstantTortuosity; D = 0: off
    tracer_diffusion: Tuple[float, float] = (0.0, 0.5)
    # tracer: mechanical dispersion -- n.D.n of the dispersion tensor at every face [n, 2*dim] (sides as volume_flux); None: off
    tracer_dispersion: Optional[np.ndarray] = None
    # slab-local spec (multi-GPU set-up without materialising the global arrays): the per-cell / per-face arrays above
    # cover only the layers [slab[0], slab[1]) of the last axis (overlap included); cells/lower/upper stay GLOBAL.
    slab: Optional[Tuple[int, int]] = None
    # box-local spec: the arrays cover the box [box[a][0], box[a][1]) of every axis (block decomposition, overlap included)
    box: Optional[Tuple[Tuple[int, int], ...]] = None

    @property
    def local_box(self):
        """per-axis (lo, hi) the per-cell arrays cover, or None if they are global"""
        return _as_box(self.cells, self.slab, self.box)

    @property
    def num_eq(self) -> int:
        return 2 if self.model == MODEL_2P else 1

    @property
    def num_cells(self) -> int:
        return int(np.prod(self.cells))

    @property
    def cells3(self) -> Tuple[int, int, int]:
        c = tuple(self.cells) + (1,) * (3 - len(self.cells))
        return c  # type: ignore


# ------------------------------------------------------------------------------------------------------
# geometry helpers (YaspGrid equidistant coordinates: x_i = lower + i*h; centres 0.5*(x_i + x_{i+1}))
# ------------------------------------------------------------------------------------------------------
def node_coords(cells, lower, upper):
    out = []
    for a in range(len(cells)):
        h = (upper[a] - lower[a]) / cells[a]
        out.append(lower[a] + np.arange(cells[a] + 1, dtype=np.float64) * h)
    return out


def axis_partition(n_layers: int, nparts: int, coord: int, overlap: int = 1):
    """Layers [lo, hi) of one axis held by the block at torus coordinate `coord` (overlap included) and its owned range
    [b0, b1).  YaspGrid's tensor-product partitioning (dune-grid torus.hh Torus::partition [DUNE-ext]): n/P layers for the
    first P - n%P blocks, one more for the others; Grid.Overlap 1 (io/grid/gridmanager_yasp.hh:129)."""
    if nparts == 1:
        return 0, n_layers, 0, n_layers
    m, rem = divmod(n_layers, nparts)
    if coord < nparts - rem:
        b0, b1 = coord * m, coord * m + m
    else:
        b0 = (nparts - rem) * m + (coord - (nparts - rem)) * (m + 1)
        b1 = b0 + m + 1
    return max(0, b0 - overlap), min(n_layers, b1 + overlap), b0, b1


def slab_partition(n_layers: int, nranks: int, rank: int, overlap: int = 1):
    """Layers [lo, hi) of the split axis held by `rank` (overlap included) and its owned range [b0, b1): the
    Yasp-style fixed-size partitioning "1 1 P" with Grid.Overlap 1 (io/grid/gridmanager_yasp.hh:129,194-203) as
    dmx_grid_structured cuts it."""
    return axis_partition(n_layers, nranks, rank, overlap)


def default_partitioning(dim: int, nranks: int):
    """Grid.Partitioning used when none is given: slabs along the last axis ("1 .. P")."""
    return tuple([1] * (dim - 1) + [nranks])


def rank_coord(part, rank: int):
    """Torus coordinate of `rank`, x fastest (dune-grid torus.hh Torus::rank_to_coord [DUNE-ext])."""
    c = []
    for p in part:
        c.append(rank % p)
        rank //= p
    return tuple(c)


def box_partition(cells, part, rank: int, overlap: int = 1):
    """Per axis (lo, hi, b0, b1) of the block of `rank` in a `part` = (px, py[, pz]) decomposition (Grid.Partitioning,
    io/grid/gridmanager_yasp.hh:194-203): the local box incl. overlap and the owned (interior) range, global indices."""
    coord = rank_coord(part, rank)
    return [axis_partition(cells[a], part[a], coord[a], overlap) for a in range(len(cells))]


def _cut(arrs, box):
    if box is None:
        return arrs
    return [a[lo:hi] for a, (lo, hi) in zip(arrs, box)]


def _as_box(cells, slab=None, box=None):
    """normalise the (slab | box) arguments to a per-axis list of (lo, hi) or None"""
    if box is not None:
        return [tuple(b) for b in box]
    if slab is not None:
        return [(0, c) for c in cells[:-1]] + [tuple(slab)]
    return None


def cell_centers(cells, lower, upper, slab=None, box=None):
    """float64[n, dim], x fastest; `slab` = (lo, hi) restricts the last axis to those layers, `box` = per-axis (lo, hi)."""
    xs = node_coords(cells, lower, upper)
    ctr = _cut([0.5 * (x[:-1] + x[1:]) for x in xs], _as_box(cells, slab, box))
    grids = np.meshgrid(*ctr, indexing="ij")
    # x fastest: flatten in Fortran order
    return np.stack([g.reshape(-1, order="F") for g in grids], axis=1)


def side_face_centers(cells, lower, upper, side, slab=None, box=None):
    """float64[nf, dim] centres of the boundary faces of `side` (lower remaining axis fastest) of the GLOBAL boundary,
    restricted to the local box in the other axes."""
    dim = len(cells)
    a = side // 2
    xs = node_coords(cells, lower, upper)
    ctr = _cut([0.5 * (x[:-1] + x[1:]) for x in xs], _as_box(cells, slab, box))
    ctr[a] = np.array([xs[a][-1] if side & 1 else xs[a][0]])
    grids = np.meshgrid(*ctr, indexing="ij")
    return np.stack([g.reshape(-1, order="F") for g in grids], axis=1)[:, :dim]


def _in_box(pos, lo, hi, eps):
    ok = np.ones(pos.shape[0], dtype=bool)
    for a in range(pos.shape[1]):
        ok &= ~((pos[:, a] < lo[a] + eps) | (pos[:, a] > hi[a] - eps))
    return ok


# ------------------------------------------------------------------------------------------------------
# std::mt19937 + Dumux::SimpleLogNormalDistribution (dumux/common/random.hh; Box-Mueller with cached pair)
# ------------------------------------------------------------------------------------------------------
class DumuxLogNormalField:
    """Replays ``std::mt19937 rand(seed)`` feeding one or more ``SimpleLogNormalDistribution`` objects.

    Each distribution keeps its own Box-Mueller cache (dumux/common/random.hh SimpleNormalDistribution::operator()),
    returns z1 first and the cached z0 on the next call; lognormal = exp(normal).
    """

    def __init__(self, seed: int = 0):
        self._rs = np.random.RandomState(seed)          # init_genrand(seed) == std::mt19937(seed)
        self._buf = np.empty(0, dtype=np.uint64)
        self._pos = 0

    def _next_u32(self) -> int:
        if self._pos >= self._buf.size:
            # RandomState.randint draws masked 32-bit outputs one per value for this range
            self._buf = self._rs.randint(0, 2 ** 32, size=1 << 16, dtype=np.uint64)
            self._pos = 0
        v = int(self._buf[self._pos])
        self._pos += 1
        return v

    def make_dist(self, mean: float, stddev: float):
        state = {"cached": None}
        eps = np.finfo(np.float64).eps

        def draw() -> float:
            if state["cached"] is not None:
                v = state["cached"]
                state["cached"] = None
                return float(np.exp(v))
            while True:
                u1 = 2.0 ** -32 * self._next_u32()
                u2 = 2.0 ** -32 * self._next_u32()
                if u1 > eps:
                    break
            magnitude = stddev * np.sqrt(-2.0 * np.log(u1))
            z0 = magnitude * np.cos(2.0 * np.pi * u2) + mean
            z1 = magnitude * np.sin(2.0 * np.pi * u2) + mean
            state["cached"] = z0
            return float(np.exp(z1))

        return draw


def lognormal_permeability(n: int, kmean: float, seed: int = 0, in_lens: Optional[np.ndarray] = None,
                           kmean_lens: Optional[float] = None) -> np.ndarray:
    """examples/1ptracer/spatialparams_1p.hh:95-104: one draw per element in element order."""
    gen = DumuxLogNormalField(seed)
    dist = gen.make_dist(np.log(kmean), np.log(kmean) * 0.1)
    dist_lens = gen.make_dist(np.log(kmean_lens), np.log(kmean_lens) * 0.1) if kmean_lens is not None else None
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        if in_lens is not None and in_lens[i]:
            out[i] = dist_lens()
        else:
            out[i] = dist()
    return out


def plane_lognormal_multiplier(cells, sigma: float, seed: int = 0, slab=None, box=None) -> np.ndarray:
    """exp(N(0, sigma)) per cell with one MT19937 stream per layer of the last axis (seeded seed*1000003 + layer), so
    that any slab / block of a decomposed grid can be generated without the global field."""
    bx = _as_box(cells, slab, box)
    lo, hi = (0, cells[-1]) if bx is None else bx[-1]
    inner = tuple(cells[:-1])
    per_layer = int(np.prod(inner))
    sl = tuple(slice(*b) for b in reversed(bx[:-1])) if bx is not None else None
    out = []
    for k in range(lo, hi):
        layer = np.exp(np.random.RandomState(seed * 1000003 + k).normal(0.0, sigma, size=per_layer))
        if sl is not None:
            layer = layer.reshape(tuple(reversed(inner)))[sl].reshape(-1)
        out.append(layer)
    return np.concatenate(out)


def fast_lognormal_multiplier(n: int, sigma: float, seed: int = 0) -> np.ndarray:
    """Vectorised heterogeneity multiplier exp(N(0, sigma)) for the large synthetic grids:
    same generator family (MT19937 seeded with init_genrand(seed)), numpy's normal transform."""
    rs = np.random.RandomState(seed)
    return np.exp(rs.normal(0.0, sigma, size=n))


# ------------------------------------------------------------------------------------------------------
# C1: test/porousmediumflow/1p/incompressible (params.input, problem.hh:41-100, spatialparams.hh:44-106)
# ------------------------------------------------------------------------------------------------------
def onep_incompressible(cells=(10, 10), lower=None, upper=None, numdiff_params=True, analytic=False) -> ProblemSpec:
    """`analytic`: DiffMethod::analytic (test_1p_incompressible_tpfa, the reference's default) instead of numeric differentiation."""
    dim = len(cells)
    lower = tuple([0.0] * dim) if lower is None else lower
    upper = tuple([1.0] * dim) if upper is None else upper
    n = int(np.prod(cells))
    ctr = cell_centers(cells, lower, upper)
    lens = _in_box(ctr, [0.2] * dim, [0.8] * dim, 1.5e-7)
    K = np.where(lens, 1e-12, 1e-10)
    bc_type, bc_values = {}, {}
    ymax = upper[dim - 1]
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side)
        y = fc[:, dim - 1]
        dirichlet = (y < 1e-6) | (y > ymax - 1e-6)
        bc_type[side] = np.where(dirichlet, BC_DIRICHLET, BC_NEUMANN).astype(np.int32)
        vals = np.zeros((fc.shape[0], 1))
        vals[dirichlet, 0] = 1.0e5 + (-1.0e5) * (y[dirichlet] - ymax)
        bc_values[side] = vals
    opt = Options(stationary=True)
    if numdiff_params:
        opt.base_eps = 0.1
        opt.privar_magnitude = (1e5, -1.0)
    if analytic:
        opt.fd_method = DIFF_ANALYTIC
    return ProblemSpec(
        name="1p_incompressible", model=MODEL_1P, dim=dim, cells=tuple(cells), lower=tuple(lower), upper=tuple(upper),
        K=K, phi=np.full(n, 0.4), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1000.0,), mu=(1e-3,), bc_type=bc_type, bc_values=bc_values, options=opt,
        initial=np.zeros((n, 1)))


# ------------------------------------------------------------------------------------------------------
# C2: test/porousmediumflow/1p/compressible/instationary (params.input, problem.hh:41-108, spatialparams.hh:44-98):
# tabulated H2O (TabulatedComponent<H2O>::init(273.15, 294.15, 10, 1e4, 1e6, 200)), T = 293.15 K, Dirichlet p = 1e5*(2 - z)
# at bottom/top, no-flow sides, initial p = 1e5, default FD step.  `lognormal=True` swaps the lens for the log-normal
# permeability field of examples/1ptracer (BASELINE config 2: "lognormal random permeability").
# ------------------------------------------------------------------------------------------------------
_H2O_TABLE = None


def onep_compressible(cells=(10, 10), lower=None, upper=None, dt=0.002, lognormal=False, seed=0) -> ProblemSpec:
    global _H2O_TABLE
    from . import iapws
    if _H2O_TABLE is None:
        _H2O_TABLE = iapws.tabulated_h2o()
    dim = len(cells)
    lower = tuple([0.0] * dim) if lower is None else lower
    upper = tuple([1.0] * dim) if upper is None else upper
    n = int(np.prod(cells))
    ctr = cell_centers(cells, lower, upper)
    lens = _in_box(ctr, [0.2] * dim, [0.8] * dim, 1.5e-7)
    if lognormal:
        K = lognormal_permeability(n, 1e-10, seed, lens, 1e-11) if n <= 4_000_000 else \
            np.where(lens, 1e-11, 1e-10) * fast_lognormal_multiplier(n, 0.5, seed)
    else:
        K = np.where(lens, 1e-12, 1e-10)
    bc_type, bc_values = {}, {}
    zmax = upper[dim - 1]
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side)
        z = fc[:, dim - 1]
        dirichlet = (z < 1e-6) | (z > zmax - 1e-6)
        bc_type[side] = np.where(dirichlet, BC_DIRICHLET, BC_NEUMANN).astype(np.int32)
        vals = np.zeros((fc.shape[0], 1))
        vals[dirichlet, 0] = 1.0e5 * (2.0 - z[dirichlet])
        bc_values[side] = vals
    return ProblemSpec(
        name="1p_compressible", model=MODEL_1P, dim=dim, cells=tuple(cells), lower=tuple(lower), upper=tuple(upper),
        K=K, phi=np.full(n, 0.4), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1000.0,), mu=(1e-3,), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=False, dt=dt), initial=np.full((n, 1), 1.0e5), fluid_table=dict(_H2O_TABLE))


# ------------------------------------------------------------------------------------------------------
# 2p lens, the reference test: test/porousmediumflow/2p/incompressible (params.input, problem.hh:50-168,
# spatialparams.hh:46-140).  2-D: y vertical.  `vertical_axis` = dim-1 always (gravity acts along -e_{dim-1}).
# ------------------------------------------------------------------------------------------------------
def twop_lens(cells=(48, 32), law="vg", lower=None, upper=None, lens_lower=None, lens_upper=None,
              dt=250.0, heterogeneity_sigma=0.0, seed=0, bc_params=None, slab=None, plane_rng=False, oilwet=False,
              analytic=False, box=None) -> ProblemSpec:
    """`slab` = (lo, hi): build only those layers of the last axis (see ProblemSpec.slab), `box` = per-axis (lo, hi): build only
    that block (ProblemSpec.box); `plane_rng`: per-layer heterogeneity streams (implied by `slab` / `box`).  `oilwet`: test_2p_incompressible_tpfa_oilwet (SpatialParams.LensIsOilWet,
    Problem.EnableGravity false): the lens keeps the outer permeability and pc-kr-Sw parameters but phase 1 wets it
    (spatialparams.hh:76,105,117-122) and the injection rate is ten times higher (problem.hh:112-113).
    `analytic`: DiffMethod::analytic (test_2p_incompressible_tpfa_analytic)."""
    dim = len(cells)
    if dim == 2:
        lower = (0.0, 0.0) if lower is None else lower
        upper = (6.0, 4.0) if upper is None else upper
        lens_lower = (1.0, 2.0) if lens_lower is None else lens_lower
        lens_upper = (4.0, 3.0) if lens_upper is None else lens_upper
    else:
        # 3-D extension: [0,6]x[0,4]x[0,4], z vertical, lens box [1,4]x[1,3]x[2,3]
        lower = (0.0, 0.0, 0.0) if lower is None else lower
        upper = (6.0, 4.0, 4.0) if upper is None else upper
        lens_lower = (1.0, 1.0, 2.0) if lens_lower is None else lens_lower
        lens_upper = (4.0, 3.0, 3.0) if lens_upper is None else lens_upper
    n = int(np.prod(cells))
    bx = _as_box(cells, slab, box)
    if bx is not None:
        n = int(np.prod([hi - lo for lo, hi in bx]))
    va = dim - 1
    ctr = cell_centers(cells, lower, upper, box=bx)
    lens = _in_box(ctr, lens_lower, lens_upper, 1.5e-7)
    K = np.where(lens, 9.05e-12, 4.6e-10) if not oilwet else np.full(n, 4.6e-10)
    if heterogeneity_sigma > 0.0:
        if bx is not None or plane_rng:
            K = K * plane_lognormal_multiplier(cells, heterogeneity_sigma, seed, box=bx)
        else:
            K = K * fast_lognormal_multiplier(n, heterogeneity_sigma, seed)
    region = lens.astype(np.int32)
    if law == "vg":
        mats = [Material(LAW_VG, (0.0037, 4.7, 0.5), swr=0.05, reg=(0.01, 0.99, 0.1, 0.9)),
                Material(LAW_VG, (0.00045, 7.3, 0.5), swr=0.18, reg=(0.01, 0.99, 0.1, 0.9))]
    else:
        bp = bc_params or ((500.0, 2.0), (2000.0, 2.0))
        mats = [Material(LAW_BC, bp[0], swr=0.05, reg=(0.01,)),
                Material(LAW_BC, bp[1], swr=0.18, reg=(0.01,))]
    if oilwet:
        mats[1] = dataclasses.replace(mats[0], wetting=1)
    rho_w, g = 1000.0, (0.0 if oilwet else -9.81)
    height = upper[va] - lower[va]
    width = upper[0] - lower[0]
    alpha = 1 + 1.5 / height
    bc_type, bc_values = {}, {}
    eps = 1e-6
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side, box=bx)
        nf = fc.shape[0]
        x, y = fc[:, 0], fc[:, va]
        left = x < lower[0] + eps
        right = x > upper[0] - eps
        dirichlet = left | right
        t = np.where(dirichlet, BC_DIRICHLET, BC_NEUMANN).astype(np.int32)
        vals = np.zeros((nf, 2))
        depth = upper[va] - y
        factor = (width * alpha + (1.0 - alpha) * x) / width
        vals[dirichlet, 0] = (1e5 - factor * rho_w * g * depth)[dirichlet]
        vals[dirichlet, 1] = 0.0
        lam = (upper[0] - x) / width
        inlet = (y > upper[va] - eps) & (0.5 < lam) & (lam < 2.0 / 3.0) & ~dirichlet
        vals[inlet, 1] = -0.04 * 10 if oilwet else -0.04
        bc_type[side], bc_values[side] = t, vals
    init = np.zeros((n, 2))
    init[:, 0] = 1e5 - rho_w * g * (upper[va] - ctr[:, va])
    return ProblemSpec(
        name=f"2p_lens_{dim}d_{law}", model=MODEL_2P, dim=dim, cells=tuple(cells), lower=tuple(lower), upper=tuple(upper),
        K=K, phi=np.full(n, 0.4), region=region, materials=mats, rho=(1000.0, 1460.0), mu=(1e-3, 5.7e-4),
        bc_type=bc_type, bc_values=bc_values, options=Options(stationary=False, dt=dt, enable_gravity=not oilwet, fd_method=DIFF_ANALYTIC if analytic else 1),
        initial=init,
        slab=slab, box=None if box is None else tuple(tuple(b) for b in box))


# ------------------------------------------------------------------------------------------------------
# test_1p_incompressible_tpfa_extrude (test/porousmediumflow/1p/incompressible/CMakeLists.txt:144-152): the 1p test with
# -Problem.ExtrusionFactor 10 -Problem.CheckIsConstantVelocity true -Problem.EnableGravity false: homogeneous K (no lens,
# spatialparams.hh:70), analytic Jacobian; the reference checks that the Darcy velocity -K dp/dy / mu is reproduced exactly
# (main.cc:165-203).  `constant_velocity_check` does that on the face volume fluxes.
# ------------------------------------------------------------------------------------------------------
def onep_extrude(cells=(10, 10), extrusion=10.0) -> ProblemSpec:
    spec = onep_incompressible(cells, analytic=True)
    n = int(np.prod(cells))
    return dataclasses.replace(spec, name="1p_extrude", K=np.full(n, 1e-10),
                               options=dataclasses.replace(spec.options, extrusion=extrusion, enable_gravity=False))


def constant_velocity_check(spec, volume_flux):
    """volume flux / (face area * extrusion factor) on every face: vertical component equal to 1e-10 * 1e5 / 1e-3 to 1e-8
    relative, horizontal component below 1e-10 (the tolerances of main.cc:196-198); returns the two deviations"""
    h = [(spec.upper[a] - spec.lower[a]) / spec.cells[a] for a in range(2)]
    v = np.asarray(volume_flux) / (np.array([h[1], h[1], h[0], h[0]]) * spec.options.extrusion)
    exact = 1e-10 * 1.0e5 / 1e-3
    dev_y = max(np.abs(v[:, 3] / exact - 1).max(), np.abs(-v[:, 2] / exact - 1).max())
    return float(dev_y), float(np.abs(v[:, :2]).max())


# ------------------------------------------------------------------------------------------------------
# test/porousmediumflow/1p/convergence/analyticsolution (test_1p_convergence_analytic_tpfa_structured: params.input with
# -Problem.C 0.0, problem.hh:60-150, spatialparams.hh:60-75): stationary incompressible 1p on [0,1]^2 with density 1 and kinematic
# viscosity 1, the permeability TENSOR K = [[1, -c/(2w) sin(wx)], [-c/(2w) sin(wx), exp(-2)(1 + c cos(wx))]], w = pi, which for the
# structured TPFA variant (c = 0) is the diagonal tensor diag(1, exp(-2)); Dirichlet values from the analytic pressure
# p = (exp(y+1) + 2 - exp(2)) sin(wx) + 10 on the whole boundary and the matching source term.  The reference runs refinements
# 0..3 of a 10 x 10 grid and accepts a mean convergence rate of the discrete L2 error >= 1.8 (convergencetest.py).
# ------------------------------------------------------------------------------------------------------
def onep_convergence_exact(x, y):
    return (np.exp(y + 1.0) + 2.0 - np.exp(2.0)) * np.sin(np.pi * x) + 10.0


def onep_convergence(cells=(10, 10)) -> ProblemSpec:
    dim = 2
    lower, upper = (0.0, 0.0), (1.0, 1.0)
    n = int(np.prod(cells))
    om = np.pi
    ctr = cell_centers(cells, lower, upper)
    x, y = ctr[:, 0], ctr[:, 1]
    q = np.zeros((n, 1))
    q[:, 0] = (-(0.0 * np.cos(om * x) + 1.0) * np.exp(y - 1.0) + 1.5 * 0.0 * np.exp(y + 1.0) * np.cos(om * x)
               + om * om * (np.exp(y + 1.0) - np.exp(2.0) + 2.0)) * np.sin(om * x)
    K = np.empty((n, 2))
    K[:, 0] = 1.0
    K[:, 1] = np.exp(-2.0)
    bc_type, bc_values = {}, {}
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side)
        bc_type[side] = np.full(fc.shape[0], BC_DIRICHLET, dtype=np.int32)
        bc_values[side] = onep_convergence_exact(fc[:, 0], fc[:, 1]).reshape(-1, 1)
    return ProblemSpec(
        name="1p_convergence", model=MODEL_1P, dim=dim, cells=tuple(cells), lower=lower, upper=upper,
        K=K, phi=np.full(n, 1.0), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1.0,), mu=(1.0,), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=True, enable_gravity=False), initial=np.zeros((n, 1)), source=q)


def onep_convergence_l2_error(spec, p):
    """main.cc:39-57: sqrt(sum_scv volume (p_h - p_exact(dofPosition))^2)"""
    ctr = cell_centers(spec.cells, spec.lower, spec.upper)
    vol = np.prod([(spec.upper[a] - spec.lower[a]) / spec.cells[a] for a in range(spec.dim)])
    d = np.asarray(p).reshape(-1) - onep_convergence_exact(ctr[:, 0], ctr[:, 1])
    return float(np.sqrt(np.sum(vol * d * d)))


# ------------------------------------------------------------------------------------------------------
# test/porousmediumflow/1p/pointsources/timeindependent (params.input, problem.hh:33-130, properties.hh:40-50):
# incompressible 1p (SimpleH2O) on a 100 x 100 YaspGrid over [-1,1]^2, K = 1e-10, porosity 0.3, no gravity, Dirichlet p = 1e5 on the
# whole boundary, a point source of 10 kg/s at the origin; one time step dt = 1 s.  The origin is a grid vertex: DuMux's point-source
# helper divides the rate equally among the elements that contain the point (common/pointsource.hh BoundingBoxTreePointSourceHelper),
# i.e. 2.5 kg/s for each of the four cells around it -- a source density q = rate / volume in those cells.
# ------------------------------------------------------------------------------------------------------
def onep_pointsource(cells=(100, 100), rate=10.0) -> ProblemSpec:
    dim = 2
    lower, upper = (-1.0, -1.0), (1.0, 1.0)
    n = int(np.prod(cells))
    bc_type, bc_values = {}, {}
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side)
        bc_type[side] = np.full(fc.shape[0], BC_DIRICHLET, dtype=np.int32)
        bc_values[side] = np.full((fc.shape[0], 1), 1.0e5)
    ctr = cell_centers(cells, lower, upper)
    h = [(upper[a] - lower[a]) / cells[a] for a in range(dim)]
    holds = (np.abs(ctr[:, 0]) <= 0.5 * h[0] * (1 + 1e-9)) & (np.abs(ctr[:, 1]) <= 0.5 * h[1] * (1 + 1e-9))      # cells containing the origin
    q = np.zeros((n, 1))
    q[holds, 0] = rate / holds.sum() / (h[0] * h[1])
    return ProblemSpec(
        name="1p_pointsource", model=MODEL_1P, dim=dim, cells=tuple(cells), lower=lower, upper=upper,
        K=np.full(n, 1e-10), phi=np.full(n, 0.3), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1000.0,), mu=(1e-3,), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=False, dt=1.0, enable_gravity=False), initial=np.full((n, 1), 1.0e5), source=q)


# ------------------------------------------------------------------------------------------------------
# test/porousmediumflow/2p/buckleyleverett (params.input, problem.hh:52-140, spatialparams.hh:40-100, properties.hh:52-82):
# pseudo-1-D displacement of the non-wetting phase by water on a 100 x 1 YaspGrid over [0,100] x [0,75] m, no gravity,
# BrooksCoreyDefault with lambda 4, entry pressure 0, Swr = Snr = 0.2, K = 1.01936799e-14, porosity 0.2, both fluids with
# density 1000 and viscosity 1e-3; Dirichlet on the left (p = 2e5, Sn = Snr), Neumann elsewhere: the non-wetting phase leaves
# through the right boundary with totalVelocity * density = 3e-7 * 1000 kg/(m^2 s), no flow at top and bottom; initial
# p = 2e5, Sn = 1 - Swr.  TimeLoop: dt0 1e3 s, tEnd 1e7 s, MaxTimeStepSize 5e5 s.
# ------------------------------------------------------------------------------------------------------
def twop_buckleyleverett(cells=(100, 1), total_velocity=3e-7, injection_pressure=2e5) -> ProblemSpec:
    dim = 2
    lower, upper = (0.0, 0.0), (100.0, 75.0)
    n = int(np.prod(cells))
    swr = snr = 0.2
    mats = [Material(LAW_BC, (0.0, 4.0), swr=swr, snr=snr, reg=(0.01,))]
    bc_type, bc_values = {}, {}
    eps = 1e-6
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side)
        nf = fc.shape[0]
        left = fc[:, 0] < lower[0] + eps
        right = fc[:, 0] > upper[0] - eps
        bc_type[side] = np.where(left, BC_DIRICHLET, BC_NEUMANN).astype(np.int32)
        vals = np.zeros((nf, 2))
        vals[left, 0] = injection_pressure
        vals[left, 1] = snr
        vals[right & ~left, 1] = total_velocity * 1000.0
        bc_values[side] = vals
    init = np.zeros((n, 2))
    init[:, 0] = injection_pressure
    init[:, 1] = 1.0 - swr
    return ProblemSpec(
        name="2p_buckleyleverett", model=MODEL_2P, dim=dim, cells=tuple(cells), lower=lower, upper=upper,
        K=np.full(n, 1.01936799e-14), phi=np.full(n, 0.2), region=np.zeros(n, dtype=np.int32), materials=mats,
        rho=(1000.0, 1000.0), mu=(1e-3, 1e-3), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=False, dt=1e3, enable_gravity=False), initial=init)


class BuckleyLeverettAnalyticSolution:
    """test/porousmediumflow/2p/buckleyleverett/analyticsolution.hh:32-185: Welge tangent construction for the shock saturation
    (Brent root of f_w'(S) - (f_w(S) - f_w(Swr))/(S - Swr)), rarefaction wave behind the shock by inverting
    v/phi f_w'(S) = x/t.  `law(which, sw)` evaluates krw (1), krn (2), dkrw/dSw (4), dkrn/dSw (5) of the material law."""

    def __init__(self, law, total_velocity=3e-7, porosity=0.2, swr=0.2, snr=0.2, mu_w=1e-3, mu_n=1e-3):
        from scipy.optimize import brentq
        self.law, self.v, self.phi, self.swr, self.snr, self.mu_w, self.mu_n = law, total_velocity, porosity, swr, snr, mu_w, mu_n
        self._brentq = brentq
        self.sw_left, self.sw_right = 1.0 - snr, swr
        eps = 1e-12
        tangent = lambda sw: (self.fw(sw) - self.fw(self.swr)) / (sw - self.swr)
        self.sw_shock = brentq(lambda sw: self.dfw(sw) - tangent(sw), self.sw_right + eps, self.sw_left - eps, xtol=1e-14)
        self.shock_speed = self.v / self.phi * self.dfw(self.sw_shock)

    def fw(self, sw):
        mw, mn = self.law(1, sw) / self.mu_w, self.law(2, sw) / self.mu_n
        return mw / (mw + mn)

    def dfw(self, sw):
        mw, mn = self.law(1, sw) / self.mu_w, self.law(2, sw) / self.mu_n
        dmw, dmn = self.law(4, sw) / self.mu_w, self.law(5, sw) / self.mu_n
        return (dmw * (mw + mn) - mw * (dmw + dmn)) / ((mw + mn) * (mw + mn))

    def saturation(self, x, time):
        if time <= 0.0:
            return self.swr
        xi = x / time
        xi_lower = self.v / self.phi * self.dfw(1.0 - self.snr)
        if xi <= xi_lower:
            return 1.0 - self.snr
        if xi < self.shock_speed:
            return self._brentq(lambda sw: self.v / self.phi * self.dfw(sw) - xi, self.sw_shock, self.sw_left, xtol=1e-14)
        return self.swr

    def check(self, spec, u, t_end, density_w=1000.0):
        """main.cc:140-205: relative errors of the wetting-phase centre of mass and total mass against the analytic profile"""
        from scipy.integrate import quad
        x_min, x_max = spec.lower[0], spec.upper[0]
        width = spec.upper[1] - spec.lower[1]
        pts = [x_min + self.shock_speed * t_end] if x_min < self.shock_speed * t_end < x_max else None
        m1 = quad(lambda x: x * density_w * self.saturation(x, t_end) * self.phi * width, x_min, x_max, points=pts, limit=400)[0]
        m0 = quad(lambda x: density_w * self.saturation(x, t_end) * self.phi * width, x_min, x_max, points=pts, limit=400)[0]
        ctr = cell_centers(spec.cells, spec.lower, spec.upper)
        nodes = node_coords(spec.cells, spec.lower, spec.upper)
        vol = np.outer(np.diff(nodes[1]), np.diff(nodes[0])).reshape(-1)
        sw = 1.0 - np.asarray(u).reshape(-1, 2)[:, 1]
        mass = density_w * self.phi * vol * sw
        com_num, com_ana = float((ctr[:, 0] * mass).sum() / mass.sum()), m1 / m0
        return abs(com_ana - com_num) / com_ana, abs(m0 - float(mass.sum())) / m0


# ------------------------------------------------------------------------------------------------------
# 1p incompressible on the log-normal field of examples/1ptracer (spatialparams_1p.hh:95-104, problem_1p.hh)
# ------------------------------------------------------------------------------------------------------
def onep_tracer_pressure(cells=(50, 50)) -> ProblemSpec:
    dim = len(cells)
    lower, upper = tuple([0.0] * dim), tuple([1.0] * dim)
    n = int(np.prod(cells))
    ctr = cell_centers(cells, lower, upper)
    lens = _in_box(ctr, [0.2] * dim, [0.8] * dim, 1.5e-7)
    K = lognormal_permeability(n, 1e-10, 0, lens, 1e-11)
    ymax = upper[dim - 1]
    bc_type, bc_values = {}, {}
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side)
        y = fc[:, dim - 1]
        dirichlet = (y < 1e-6) | (y > ymax - 1e-6)
        bc_type[side] = np.where(dirichlet, BC_DIRICHLET, BC_NEUMANN).astype(np.int32)
        vals = np.zeros((fc.shape[0], 1))
        vals[dirichlet, 0] = 1.0e5 * (1.1 - y[dirichlet] * 0.1)       # problem_1p.hh:85-95: 1.1 bar bottom, 1 bar top
        bc_values[side] = vals
    # The example assembles this LINEAR problem with DiffMethod::analytic (examples/1ptracer/main.cc:112); a large FD
    # step (the numdiff test's BaseEpsilon 0.1 x PriVarMagnitude 1e5) reproduces the analytic Jacobian to rounding.
    return ProblemSpec(
        name="1ptracer_pressure", model=MODEL_1P, dim=dim, cells=tuple(cells), lower=lower, upper=upper,
        K=K, phi=np.full(n, 0.2), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1000.0,), mu=(1e-3,), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=True, base_eps=0.1, privar_magnitude=(1e5, -1.0)), initial=np.zeros((n, 1)))


# ------------------------------------------------------------------------------------------------------
# C5: tracer transport on the velocity field of the 1p problem above (examples/1ptracer: problem_tracer.hh:60-145,
# spatialparams_tracer.hh:40-110, properties_tracer.hh:60-91, params.input).  Mass fractions (UseMoles = false), one
# component, D = 0, porosity 0.2, fluid density 1000; all boundaries Neumann: outflow volumeFlux*X*rho/area at the top,
# zero elsewhere; initial X = 1e-9 * M_tracer / M_fluid (0.300 / 18.0) below y = 0.1.
# ------------------------------------------------------------------------------------------------------
def tracer_transport(cells, volume_flux, dt=10.0, implicit=False, box=None) -> ProblemSpec:
    """`box` = per-axis (lo, hi): build only that block (ProblemSpec.box); `volume_flux` then covers the block too"""
    dim = len(cells)
    lower, upper = tuple([0.0] * dim), tuple([1.0] * dim)
    bx = _as_box(cells, None, box)
    n = int(np.prod(cells)) if bx is None else int(np.prod([hi - lo for lo, hi in bx]))
    ctr = cell_centers(cells, lower, upper, box=bx)
    zmax = upper[dim - 1]
    bc_type, bc_values = {}, {}
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side, box=bx)
        top = fc[:, dim - 1] > zmax - 1e-6
        bc_type[side] = np.where(top, BC_OUTFLOW, BC_NEUMANN).astype(np.int32)
        bc_values[side] = np.zeros((fc.shape[0], 1))
    init = np.zeros((n, 1))
    init[ctr[:, dim - 1] < 0.1 + 1e-6, 0] = 1e-9 * 0.300 / 18.0
    vf = np.ascontiguousarray(volume_flux, dtype=np.float64).reshape(n, 2 * dim)
    return ProblemSpec(
        name="tracer", model=MODEL_TRACER, dim=dim, cells=tuple(cells), lower=lower, upper=upper,
        K=np.ones(n), phi=np.full(n, 0.2), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1000.0,), mu=(1e-3,), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=False, dt=dt, enable_gravity=False), initial=init, volume_flux=vf, implicit=implicit,
        box=None if box is None else tuple(tuple(b) for b in box))


def onep_tracer_pressure_large(cells, box=None, sigma=0.5, seed=0) -> ProblemSpec:
    """The stationary 1p problem of examples/1ptracer at benchmark size (BASELINE config 5): same boundary conditions and lens as
    onep_tracer_pressure, K = {1e-10, lens 1e-11} x exp(N(0, sigma)) drawn per layer (plane_lognormal_multiplier: the field
    does not depend on the decomposition; the mt19937 replay of the 50 x 50 example is a Python loop).  `box`: one block."""
    dim = len(cells)
    lower, upper = tuple([0.0] * dim), tuple([1.0] * dim)
    bx = _as_box(cells, None, box)
    n = int(np.prod(cells)) if bx is None else int(np.prod([hi - lo for lo, hi in bx]))
    ctr = cell_centers(cells, lower, upper, box=bx)
    lens = _in_box(ctr, [0.2] * dim, [0.8] * dim, 1.5e-7)
    K = np.where(lens, 1e-11, 1e-10) * plane_lognormal_multiplier(cells, sigma, seed, box=bx)
    ymax = upper[dim - 1]
    bc_type, bc_values = {}, {}
    for side in range(2 * dim):
        fc = side_face_centers(cells, lower, upper, side, box=bx)
        y = fc[:, dim - 1]
        dirichlet = (y < 1e-6) | (y > ymax - 1e-6)
        bc_type[side] = np.where(dirichlet, BC_DIRICHLET, BC_NEUMANN).astype(np.int32)
        vals = np.zeros((fc.shape[0], 1))
        vals[dirichlet, 0] = 1.0e5 * (1.1 - y[dirichlet] * 0.1)
        bc_values[side] = vals
    return ProblemSpec(
        name="1ptracer_pressure_large", model=MODEL_1P, dim=dim, cells=tuple(cells), lower=lower, upper=upper,
        K=K, phi=np.full(n, 0.2), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1000.0,), mu=(1e-3,), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=True, base_eps=0.1, privar_magnitude=(1e5, -1.0)), initial=np.zeros((n, 1)),
        box=None if box is None else tuple(tuple(b) for b in box))


# ------------------------------------------------------------------------------------------------------
# test/porousmediumflow/tracer/constvel (params.input, problem.hh:60-120, spatialparams.hh:40-100): analytic divergence-free
# velocity field sampled at the face centres (volumeFlux = v(ipGlobal) . n * area), porosity 0.2, fluid density 1000, all
# boundaries no-flow Neumann, initial band 0.4 <= y <= 0.6 with X = 1e-9 * M_tracer / M_fluid (0.3 / 18), dt = 1e4 s to 1e6 s.
# The test carries two decoupled components: D = 1e-8 (Problem.D) for the first, D2 = 0 (pure advection) for the second; one
# spec describes one of them (`D`).
# ------------------------------------------------------------------------------------------------------
def scheidegger_normal_entry(vel, axis, alpha_l, alpha_t):
    """n.D.n of Scheidegger's dispersion tensor D = (aL - aT) v v^T / |v| + aT |v| I (dispersiontensors/scheidegger.hh:152-176) for
    the face normal e_axis; `vel` = velocity vectors [m, dim] at the face centres"""
    vel = np.asarray(vel, dtype=np.float64)
    vnorm = np.sqrt(np.sum(vel * vel, axis=1))
    vv = vel[:, axis] * vel[:, axis]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(vnorm < 1e-20, 0.0, vv / vnorm)
    return t * (alpha_l - alpha_t) + vnorm * alpha_t


def tracer_constvel(cells=(50, 50), dt=1.0e4, implicit=False, D=0.0, alpha_l=0.0, alpha_t=0.0) -> ProblemSpec:
    dim = 2
    lower, upper = (0.0, 0.0), (1.0, 1.0)
    n = int(np.prod(cells))
    ctr = cell_centers(cells, lower, upper)
    xn = node_coords(cells, lower, upper)
    hx, hy = np.diff(xn[0]), np.diff(xn[1])
    i = np.arange(n) % cells[0]
    j = np.arange(n) // cells[0]

    def vel(x, y):
        vx = 1e-5 * (x * x * (1.0 - x) * (1.0 - x) * (2.0 * y - 6.0 * y * y + 4.0 * y * y * y))
        vy = 1e-5 * (-1.0 * y * y * (1.0 - y) * (1.0 - y) * (2.0 * x - 6.0 * x * x + 4.0 * x * x * x))
        return vx, vy

    vf = np.zeros((n, 4))
    yc, xc = ctr[:, 1], ctr[:, 0]
    vf[:, 0] = -vel(xn[0][i], yc)[0] * hy[j]            # -x face: outer normal (-1, 0), area = hy
    vf[:, 1] = vel(xn[0][i + 1], yc)[0] * hy[j]
    vf[:, 2] = -vel(xc, xn[1][j])[1] * hx[i]
    vf[:, 3] = vel(xc, xn[1][j + 1])[1] * hx[i]
    disp = None
    if alpha_l != 0.0 or alpha_t != 0.0:
        # test_tracer_implicit_dispersion_tpfa (-Problem.AlphaL 0.02 -Problem.AlphaT 0.008): Scheidegger's tensor from the analytic
        # velocity at the face centres (spatialparams.hh:75-87,104-107)
        disp = np.zeros((n, 4))
        faces = [(xn[0][i], yc, 0), (xn[0][i + 1], yc, 0), (xc, xn[1][j], 1), (xc, xn[1][j + 1], 1)]
        for side, (fx, fy, axis) in enumerate(faces):
            v = np.stack(vel(fx, fy), axis=1)
            disp[:, side] = scheidegger_normal_entry(v, axis, alpha_l, alpha_t)
    bc_type, bc_values = {}, {}
    for side in range(4):
        nf = cells[1] if side < 2 else cells[0]
        bc_type[side] = np.full(nf, BC_NEUMANN, dtype=np.int32)
        bc_values[side] = np.zeros((nf, 1))
    init = np.zeros((n, 1))
    band = (ctr[:, 1] > 0.4 - 1e-6) & (ctr[:, 1] < 0.6 + 1e-6)
    init[band, 0] = 1e-9 * 0.300 / 18.0
    return ProblemSpec(
        name="tracer_constvel", model=MODEL_TRACER, dim=dim, cells=tuple(cells), lower=lower, upper=upper,
        K=np.ones(n), phi=np.full(n, 0.2), region=np.zeros(n, dtype=np.int32), materials=[],
        rho=(1000.0,), mu=(1e-3,), bc_type=bc_type, bc_values=bc_values,
        options=Options(stationary=False, dt=dt, enable_gravity=False), initial=init, volume_flux=vf, implicit=implicit,
        tracer_diffusion=(D, 0.5), tracer_dispersion=disp)
