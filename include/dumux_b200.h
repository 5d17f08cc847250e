/*
 * dumux_b200.h -- C ABI of the B200-native Newton-step engine (libdumux_b200.so).
 *
 * This is the drop-in boundary for DuMux's Newton-step hot path (cell-centred TPFA assembly + Krylov solve).
 * Every entry point names the reference interface it replaces (paths relative to the DuMux source tree).
 * Plain pointers and sizes only; no C++ or torch types.  All functions return 0 on success, a positive
 * DMX_STATUS_* for recoverable numerical conditions (what DuMux turns into Dumux::NumericalProblem,
 * dumux/nonlinear/newtonsolver.hh:514-523) and a negative value for CUDA/NCCL/usage errors; the message is
 * available from dmx_last_error().  One dmx_ctx per GPU; a ctx is not thread-safe; all work is stream-ordered
 * on the ctx's stream.  There is no CPU fallback: every compute entry point runs CUDA kernels or fails.
 *
 * Data layout at the boundary (what Dune::BCRSMatrix / Dune::BlockVector store):
 *   vectors   double[n*b], block i at [i*b, i*b+b)                      (BlockVector<FieldVector<double,b>>)
 *   matrix    rowptr int32[n+1], colidx int32[nnzb] ascending per row, values double[nnzb*b*b] row-major blocks
 *   cells     numbered x fastest (YaspGrid); boundary faces of a side numbered lower remaining axis fastest
 *   sides     0:-x 1:+x 2:-y 3:+y 4:-z 5:+z (intersection.indexInInside())
 */
#ifndef DUMUX_B200_H
#define DUMUX_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dmx_ctx dmx_ctx;

enum { DMX_MODEL_1P = 1, DMX_MODEL_2P = 2, DMX_MODEL_TRACER = 3 };
enum { DMX_LAW_BROOKSCOREY = 0, DMX_LAW_VANGENUCHTEN = 1 };
/* DMX_BC_OUTFLOW (tracer model only): the solution-dependent Neumann flux volumeFlux * X * rho / area of
   examples/1ptracer/problem_tracer.hh:92-115 */
enum { DMX_BC_NEUMANN = 0, DMX_BC_DIRICHLET = 1, DMX_BC_NONE = 2, DMX_BC_OUTFLOW = 3 };
/* dmx_options.fd_method value selecting DiffMethod::analytic instead of numeric differentiation: CCLocalAssembler<analytic,
   implicit> (assembly/cclocalassembler.hh:490-600) with OnePIncompressibleLocalResidual (porousmediumflow/1p/
   incompressiblelocalresidual.hh:76-123,204-221).  Incompressible 1p model only (constant density and viscosity). */
enum { DMX_DIFF_ANALYTIC = 100 };
/* SSOR = Dune::SeqSSOR(1 iteration, relaxation 1): the preconditioner of SSORCGIstlSolver / SSORBiCGSTABIstlSolver
   (linear/istlsolvers.hh:686-714) */
/* PARMT_*: DuMux's own multi-threaded smoothers Dumux::ParMTJac / ParMTSOR / ParMTSSOR ("par_mt_jac", "par_mt_sor", "par_mt_ssor";
   dumux/linear/preconditioners.hh:330-400, 489-620): Jacobi, and (S)SOR colour by colour with the greedy colouring of
   computeColorsForMatrixSweep_ (:408-440); iterations / relaxation from dmx_set_preconditioner_params */
/* AMG: aggregation multigrid V-cycle (AMGBiCGSTABIstlSolver / AMGCGIstlSolver, linear/istlsolvers.hh:716-757), see dmx_amg_params */
enum { DMX_PRECOND_ILU0 = 0, DMX_PRECOND_BLOCKJACOBI = 1, DMX_PRECOND_SSOR = 2, DMX_PRECOND_PARMT_JAC = 3, DMX_PRECOND_PARMT_SOR = 4,
       DMX_PRECOND_PARMT_SSOR = 5, DMX_PRECOND_AMG = 6 };
/* Krylov method behind dmx_linear_solve / dmx_newton_*: ILUBiCGSTABIstlSolver (linear/istlsolvers.hh:636-642, default) or
   ILURestartedGMResIstlSolver (:660-667) */
enum { DMX_SOLVER_BICGSTAB = 0, DMX_SOLVER_RESTARTED_GMRES = 1, DMX_SOLVER_CG = 2 };   /* CG: Dune::CGSolver (SSORCGIstlSolver :701-714) */
enum { DMX_STATUS_OK = 0, DMX_STATUS_NOT_CONVERGED = 1, DMX_STATUS_BREAKDOWN = 2, DMX_STATUS_NONFINITE = 3 };
/* device-resident vectors of a ctx */
enum {
    DMX_VEC_CUR = 0,      /* curSol: current Newton iterate                      */
    DMX_VEC_PREV = 1,     /* prevSol: previous time level (fvassembler.hh:650)   */
    DMX_VEC_RESIDUAL = 2, /* residual r                                          */
    DMX_VEC_DELTA = 3,    /* deltaU, the linear-solve unknown                    */
    DMX_VEC_ULAST = 4,    /* uLastIter (newtonsolver.hh:985)                     */
    DMX_VEC_WORK0 = 5,    /* scratch for tests / SpMV input                      */
    DMX_VEC_WORK1 = 6
};

/* Runtime parameters that change arithmetic (dumux/common/parameters.cc:231-260, assembly/numericepsilon.hh:37-53) */
typedef struct {
    int    enable_gravity;       /* Problem.EnableGravity (default 1)                                       */
    double gravity;              /* 9.81, along -e_{dim-1} (common/fvspatialparams.hh:48-52)                */
    double upwind_weight;        /* Flux.UpwindWeight (default 1.0; flux/upwindscheme.hh:42)                */
    int    fd_method;            /* Assembly.NumericDifferenceMethod: 1 forward, 0 central, -1 backward, 5  */
    double base_eps;             /* Assembly.NumericDifference.BaseEpsilon (default 1e-10)                  */
    double privar_magnitude[2];  /* Assembly.NumericDifference.PriVarMagnitude (<=0: unset)                 */
    int    stationary;           /* FVAssembler stationary ctor (fvassembler.hh:131): no storage term       */
    double dt;                   /* timeLoop->timeStepSize() used by the storage term                       */
    double extrusion;            /* constant extrusion factor                                               */
} dmx_options;

/* Newton parameters (dumux/nonlinear/newtonsolver.hh:1213-1247) + linear solver (linearsolverparameters.hh:56-73) */
typedef struct {
    double max_relative_shift;   /* Newton.MaxRelativeShift 1e-8 */
    int    min_steps;            /* Newton.MinSteps 2            */
    int    max_steps;            /* Newton.MaxSteps 18           */
    double lin_reduction;        /* LinearSolver.ResidualReduction 1e-6 (newtonsolver.hh:232) */
    int    lin_maxit;            /* LinearSolver.MaxIterations 250 */
    int    preconditioner;       /* DMX_PRECOND_* */
    /* update strategy and convergence criteria (newtonsolver.hh:1213-1232, 543-566, 657-701, 1154-1178) */
    int    use_line_search;               /* Newton.UseLineSearch 0: halve the update until the residual norm decreases */
    double line_search_min_relaxation;    /* Newton.LineSearchMinRelaxationFactor 0.125 */
    int    enable_shift_criterion;        /* Newton.EnableShiftCriterion 1 */
    int    enable_residual_criterion;     /* Newton.EnableResidualCriterion 0 */
    int    enable_absolute_residual_criterion;   /* Newton.EnableAbsoluteResidualCriterion 0 (implies the residual criterion) */
    int    satisfy_residual_and_shift;    /* Newton.SatisfyResidualAndShiftCriterion 0 */
    double residual_reduction;            /* Newton.ResidualReduction 1e-5 */
    double max_absolute_residual;         /* Newton.MaxAbsoluteResidual 1e-5 */
} dmx_newton_params;

typedef struct {
    int    newton_iterations;
    int    converged;
    int    linear_iterations_total;
    double last_shift;
    double t_assemble, t_solve, t_update;   /* seconds, CUDA-event timed; buckets of newtonsolver.hh:950-955 */
    int    linear_iterations[64];
    double shifts[64];
    double last_reduction;       /* residual norm / initial residual norm (line search or residual criterion only) */
    double last_residual_norm;
    double relaxation[64];       /* line search: the accepted lambda of every iteration */
} dmx_newton_report;

/* Parameters of the AMG preconditioner, named after dune-istl's (Dune::Amg::Parameters / Dune::AMGCreator keys) with its
   defaults: V-cycle, preSteps = postSteps = 2, prolongationDampingFactor 1.6, smoother SeqSSOR (1 iteration, relaxation 1).
   Aggregates are 2x2x2 cell blocks of the structured grid (dune's default aggregate size is 4..8); the hierarchy ends at
   <= coarsest_cells cells where coarsest_steps smoothing steps replace dune's direct coarse solve.  csrc/amg.cu. */
typedef struct {
    int    pre_steps;             /* preSteps  2 */
    int    post_steps;            /* postSteps 2 */
    double prolongation_damping;  /* prolongationDampingFactor 1.6 */
    int    smoother;              /* DMX_PRECOND_SSOR (default, "ssor"), DMX_PRECOND_ILU0 ("ilu"), or DuMux's multithreaded smoothers
                                     DMX_PRECOND_PARMT_JAC / _SOR / _SSOR (Dune::Amg::AMG<LOP, Vector, Dumux::ParMTSSOR<...>>,
                                     test/linear/test_parallel_amg_smoothers.cc:25-40, linear/stokes_solver.hh:261) */
    int    coarsest_cells;        /* stop coarsening at <= this many cells (default 8) */
    int    coarsest_steps;        /* smoothing steps on the coarsest level (default 8) */
    int    max_levels;            /* maxLevel 15 */
    int    smoother_iterations;   /* SmootherArgs::iterations (default 1); honoured by the ParMT smoothers, must be 1 for ssor / ilu */
    double smoother_relaxation;   /* SmootherArgs::relaxationFactor (default 1.0); honoured by the ParMT smoothers, must be 1 for ssor / ilu */
} dmx_amg_params;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
void dmx_default_options(dmx_options* o);
void dmx_default_newton_params(dmx_newton_params* p);
void dmx_default_amg_params(dmx_amg_params* p);
int  dmx_set_amg_params(dmx_ctx* ctx, const dmx_amg_params* p);
/* number of levels of the AMG hierarchy (0 before the first set-up) and the cells per axis of one level */
int  dmx_amg_levels(dmx_ctx* ctx);
int  dmx_amg_level_cells(dmx_ctx* ctx, int level, int* cells);
/* developer / test access: the BCRS values of a coarse level's Galerkin matrix (host buffer, nnz blocks of that level) */
long long dmx_amg_level_nnz_blocks(dmx_ctx* ctx, int level);
int  dmx_amg_level_matrix(dmx_ctx* ctx, int level, double* values);
int  dmx_create(dmx_ctx** out, int device);
/* One ctx per rank of a slab-decomposed run; nccl_unique_id = the 128-byte ncclUniqueId from rank 0.
   Replaces the MPI communicator DuMux gets from Dune::MPIHelper / gridView.comm() (linear/istlsolvers.hh:192). */
int  dmx_create_distributed(dmx_ctx** out, int device, const void* nccl_unique_id, int rank, int nranks);
int  dmx_get_nccl_unique_id(void* out128);
int  dmx_destroy(dmx_ctx* ctx);
const char* dmx_last_error(const dmx_ctx* ctx);
const char* dmx_version(void);

/* ---- grid + pattern (GridManager<YaspGrid>, io/grid/gridmanager_yasp.hh:84-135; CCTpfaFVGridGeometry::update_,
        discretization/cellcentered/tpfa/fvgridgeometry.hh:216-345; getJacobianPattern, assembly/jacobianpattern.hh:27-52) ---- */
/* cells/lower/upper describe the GLOBAL grid.  In a distributed ctx the grid is block-decomposed with overlap 1
   (Grid.Overlap default, gridmanager_yasp.hh:129): dmx_set_partitioning gives the ranks per axis (Grid.Partitioning
   "px py pz" -> Yasp::FixedSizePartitioning, gridmanager_yasp.hh:194-203; product = nranks; rank -> block with x fastest);
   without it the grid is cut into slabs along the last axis ("1 1 P").  Call before dmx_grid_*; NULL restores the default. */
int  dmx_set_partitioning(dmx_ctx* ctx, const int* ranks_per_axis);
int  dmx_grid_structured(dmx_ctx* ctx, int model, int dim, const int* cells, const double* lower, const double* upper);
int  dmx_grid_tensor(dmx_ctx* ctx, int model, int dim, const int* cells, const double* x, const double* y, const double* z);
/* local box (incl. overlap) of this rank: cells[3], offset[3] in the global index space; owned range of the last grid axis */
int  dmx_local_box(const dmx_ctx* ctx, int* cells, int* offset, int* owned_begin, int* owned_end);
/* the same with the owned range [owned_begin[a], owned_end[a]) of every axis (local indices), the ranks per axis and this
   rank's block coordinate (either may be NULL): the interior / overlap partition types of the YaspGrid piece
   (linear/parallelhelpers.hh:434-458,485-497 turn them into owner / copy attributes) */
int  dmx_local_box3(const dmx_ctx* ctx, int* cells, int* offset, int* owned_begin, int* owned_end, int* ranks_per_axis, int* coord);
int  dmx_num_cells(const dmx_ctx* ctx);
int  dmx_num_eq(const dmx_ctx* ctx);
long long dmx_nnz_blocks(const dmx_ctx* ctx);
int  dmx_pattern(const dmx_ctx* ctx, int* rowptr, int* colidx);
/* generic BCRS pattern instead of a grid (the ILUBiCGSTABIstlSolver::solve(A,x,b) entry, linear/istlsolvers.hh:273) */
int  dmx_bcrs_pattern(dmx_ctx* ctx, int n, int b, const int* rowptr, const int* colidx);

/* ---- Problem / SpatialParams sampled into flat arrays (common/fvproblem.hh:126-283,
        porousmediumflow/fvspatialparams.hh:83-99, common/fvporousmediumspatialparams.hh:75-118).  LOCAL arrays. ---- */
int  dmx_set_options(dmx_ctx* ctx, const dmx_options* o);
int  dmx_set_cell_fields(dmx_ctx* ctx, const double* permeability, const double* porosity, const int* region);
/* Diagonal permeability tensor K = diag(Kx, Ky, Kz) per cell (SpatialParams::permeability returning a FieldMatrix whose
   off-diagonal entries vanish, e.g. test/porousmediumflow/1p/convergence/analyticsolution with Problem.C = 0): on an axis-aligned
   grid the TPFA transmissibility of a face with normal e_a needs n.K.n = K_aa and the gravity term n.K.g = K_aa g_a only.
   Arrays of length num_cells per grid axis (LOCAL box), NULL for axes the grid does not have; NULL for all = back to the scalar
   field of dmx_set_cell_fields.  Full tensors (off-diagonal entries) are not supported by the TPFA kernels. */
int  dmx_set_permeability_diagonal(dmx_ctx* ctx, const double* kx, const double* ky, const double* kz);
int  dmx_set_source(dmx_ctx* ctx, const double* q);
/* BC: params {pcEntry, lambda}, reg {pcLowSwe}; VG: params {alpha, n, l}, reg {pcLowSwe, pcHighSwe, krnLowSwe, krwHighSwe} */
int  dmx_set_material(dmx_ctx* ctx, int region, int law, const double* params, double swr, double snr,
                      int regularize, const double* reg);
/* FVSpatialParams::wettingPhase of a region (porousmediumflow/2p/volumevariables.hh:87-96,132-152; e.g. the oil-wet lens of
   test_2p_incompressible_tpfa_oilwet, test/porousmediumflow/2p/incompressible/spatialparams.hh:117-122): 0 (default) = phase 0
   wets, 1 = phase 1 wets.  Call after dmx_set_material(region, ...). */
int  dmx_set_wetting_phase(dmx_ctx* ctx, int region, int phase);
int  dmx_set_fluids(dmx_ctx* ctx, const double* density, const double* viscosity);
int  dmx_set_fluid_table(dmx_ctx* ctx, int nT, int nP, double Tmin, double Tmax, const double* pmin, const double* pmax,
                         const double* density, const double* viscosity, double temperature);
/* ---- tracer transport on a frozen velocity field (DMX_MODEL_TRACER; examples/1ptracer) --------------------- */
/* 1p ctx: volume fluxes over every face of every cell from the pressure in CUR (main.cc:162-199, upwind term = mobility;
   Neumann boundary faces stay 0): out[n * 2*dim], sides -x,+x,-y,+y,-z,+z as seen from the cell.  Host buffer. */
int  dmx_volume_flux(dmx_ctx* ctx, double* out);
/* tracer ctx: TracerTestSpatialParams::setVolumeFlux (spatialparams_tracer.hh:91-103), same layout.  Host buffer. */
int  dmx_set_volume_flux(dmx_ctx* ctx, const double* volume_flux);
/* tracer ctx: FVAssembler<TracerTypeTag, DiffMethod::analytic, implicit> (main.cc:236): 0 explicit Euler, 1 implicit.
   Fluid density = density[0] of dmx_set_fluids, porosity from dmx_set_cell_fields, dt / extrusion / upwind weight from
   dmx_options.  dmx_assemble / dmx_newton_step then run TracerLocalResidual (porousmediumflow/tracer/localresidual.hh). */
int  dmx_set_tracer(dmx_ctx* ctx, int implicit);
/* tracer ctx: Fick's law (flux/cctpfa/fickslaw.hh) with DiffusivityConstantTortuosity (material/fluidmatrixinteractions/
   diffusivityconstanttortuosity.hh:55-63): D = FluidSystem::binaryDiffusionCoefficient (constant), tortuosity =
   SpatialParams.Tortuosity (default 0.5); mass fractions, mass-averaged reference system.  D = 0 (default): no diffusion. */
/* Mechanical dispersion of the tracer model (EnableCompositionalDispersion; flux/cctpfa/dispersionflux.hh:66-113 with e.g.
   ScheideggersDispersionTensor, material/fluidmatrixinteractions/dispersiontensors/scheidegger.hh:44-176): with a stationary velocity
   field the dispersion tensor at a face does not depend on the solution, so the adapter samples it once like the volume fluxes:
   disp[cell][side] = n.D.n (sides -x,+x,-y,+y,-z,+z; LOCAL box) -- the only entry a TPFA transmissibility of an axis-aligned face
   uses.  The flux rho * tij(D) * (X_I - X_J) enters the residual; like the reference's analytic tracer Jacobian
   (tracer/localresidual.hh:237-291) the Jacobian has no dispersion derivative.  NULL = off. */
int  dmx_set_tracer_dispersion(dmx_ctx* ctx, const double* disp);
int  dmx_set_tracer_diffusion(dmx_ctx* ctx, double D, double tortuosity);
int  dmx_side_faces(const dmx_ctx* ctx, int side);
int  dmx_set_boundary(dmx_ctx* ctx, int side, const int* type, const double* values);

/* ---- device-resident vectors ---------------------------------------------------------------------------- */
int  dmx_vec_upload(dmx_ctx* ctx, int vec, const double* host);
int  dmx_vec_download(dmx_ctx* ctx, int vec, double* host);
int  dmx_vec_copy(dmx_ctx* ctx, int dst, int src);
int  dmx_jacobian_upload(dmx_ctx* ctx, const double* values);
int  dmx_jacobian_download(dmx_ctx* ctx, double* values);
void* dmx_vec_device_ptr(dmx_ctx* ctx, int vec);
void* dmx_jacobian_device_ptr(dmx_ctx* ctx);

/* ---- hot path -------------------------------------------------------------------------------------------- */
/* FVAssembler::assembleJacobianAndResidual / assembleResidual (assembly/fvassembler.hh:179-207,240-267):
   reads CUR (+PREV), writes RESIDUAL (+ Jacobian values).  Returns DMX_STATUS_NONFINITE if the residual is not finite. */
int  dmx_assemble(dmx_ctx* ctx, int with_jacobian);
/* host-buffer form (the call a DuMux main makes): uploads curSol (+prevSol), assembles, downloads what is non-NULL */
int  dmx_assemble_host(dmx_ctx* ctx, const double* cur, const double* prev, double* residual, double* jacobian);
/* IstlIterativeLinearSolver::solve(A, x, b) (linear/istlsolvers.hh:273,457-464): fresh preconditioner + BiCGSTAB,
   Jacobian * DELTA = RESIDUAL, DELTA zeroed first (newtonsolver.hh:1032). */
/* selects the Krylov method (DMX_SOLVER_*); restart = LinearSolver.GMResRestart (<= 0: 10, linearsolverparameters.hh:115,138).
   For GMRes `achieved_reduction` refers to the preconditioned defect, as in Dune::RestartedGMResSolver. */
int  dmx_set_linear_solver(dmx_ctx* ctx, int solver, int restart);
int  dmx_linear_solve(dmx_ctx* ctx, double reduction, int maxit, int preconditioner, int* iterations, double* achieved_reduction);
/* host-buffer form: A values, x (in: initial guess, out: solution), b */
int  dmx_linear_solve_host(dmx_ctx* ctx, const double* values, double* x, const double* b, double reduction, int maxit,
                           int preconditioner, int* iterations, double* achieved_reduction);
/* IstlIterativeLinearSolver::norm (linear/istlsolvers.hh:306-337): owner-masked 2-norm, all-reduced */
int  dmx_norm2(dmx_ctx* ctx, int vec, double* out);
/* NewtonSolver::newtonUpdate + newtonComputeShift_ (nonlinear/newtonsolver.hh:543-557,1138-1144):
   CUR = ULAST - DELTA; shift = max relative shift (all-reduced max) */
int  dmx_newton_update(dmx_ctx* ctx, double* shift);
/* NewtonSolver::solveImpl_ (nonlinear/newtonsolver.hh:976-1072) on the device-resident state */
int  dmx_newton_solve(dmx_ctx* ctx, const dmx_newton_params* params, dmx_newton_report* report);
/* host-buffer Newton solve: uploads u and prev, solves, downloads u */
int  dmx_newton_solve_host(dmx_ctx* ctx, double* u, const double* prev, const dmx_newton_params* params, dmx_newton_report* report);
/* one Newton iteration (assemble + solve + update) on device state; used by bench.py */
int  dmx_newton_step(dmx_ctx* ctx, const dmx_newton_params* params, int* linear_iterations, double* shift, float* ms_assemble,
                     float* ms_solve, float* ms_update);
/* host-buffer Newton iteration (what one pass of the loop body of NewtonSolver::solveImpl_, newtonsolver.hh:998-1062,
   costs a DuMux main that keeps curSol on the host): uploads u, assembles, solves, updates, downloads the new u */
int  dmx_newton_step_host(dmx_ctx* ctx, double* u, const dmx_newton_params* params, int* linear_iterations, double* shift);
/* gridVariables->advanceTimeStep(): PREV = CUR;  resetTimeStep: CUR = PREV (discretization/fvgridvariables.hh:99-117) */
int  dmx_advance_timestep(dmx_ctx* ctx);
int  dmx_reset_timestep(dmx_ctx* ctx);

/* ---- kernel-level entry points (parity tests and roofline measurement) ----------------------------------- */
int  dmx_spmv(dmx_ctx* ctx, int x_vec, int y_vec);                       /* y = J x (BCRSMatrix::mv) */
int  dmx_ilu0_factor(dmx_ctx* ctx);                                      /* ILU copy of J, in place */
int  dmx_ilu0_apply(dmx_ctx* ctx, int d_vec, int v_vec);                 /* v = (LU)^-1 d */
int  dmx_ilu0_download(dmx_ctx* ctx, double* values);
/* Output fields of the solution in CUR as VtkOutputModule collects them from the volume variables (io/vtkoutputmodule.hh):
   2p: TwoPIOFields (porousmediumflow/2p/iofields.hh:31-50) S_0, p_0, rho_0, mob_0, S_1, p_1, rho_1, mob_1, pc, porosity;
   1p / tracer: the primary variable (1p/iofields.hh:30-33).  out: host buffer [dmx_num_output_fields][num_cells]. */
int  dmx_num_output_fields(const dmx_ctx* ctx);
int  dmx_output_fields(dmx_ctx* ctx, double* out);
/* v = SeqSSOR(J)(d) from v = 0: one forward + one backward block Gauss-Seidel sweep (dune-istl gsetc.hh bsorf/bsorb, w = 1) */
int  dmx_ssor_apply(dmx_ctx* ctx, int d_vec, int v_vec);
/* LinearSolver.PreconditionerIterations / PreconditionerRelaxation (linear/linearsolverparameters.hh:61-62,121-122; defaults 1 / 1.0):
   honoured by the DMX_PRECOND_PARMT_* smoothers */
int  dmx_set_preconditioner_params(dmx_ctx* ctx, int iterations, double relaxation);
/* v = M^-1 d for any DMX_PRECOND_* on this rank's matrix (sets the preconditioner up first; no halo exchange) */
int  dmx_precond_apply(dmx_ctx* ctx, int preconditioner, int d_vec, int v_vec);
int  dmx_dot(dmx_ctx* ctx, int a_vec, int b_vec, double* out);
int  dmx_halo_exchange(dmx_ctx* ctx, int vec);                           /* copyOwnerToAll */
/* average device time in ms of `reps` back-to-back launches of one kernel, CUDA-event timed on the ctx stream.
   which: 0 assembly (residual+Jacobian), 1 SpMV, 2 ILU0 apply, 3 ILU0 factor */
int  dmx_time_kernel(dmx_ctx* ctx, int which, int reps, float* ms_avg);
/* developer diagnostic: clock64 timeline of the structured ILU sweeps (enabled by DMX_SK_TRACE=1 at grid set-up):
   out[kernel 2][tile 2][chunk 64][stamp 24] */
int  dmx_debug_sweep_trace(dmx_ctx* ctx, long long* out6144);
int  dmx_kernel_launch_count(const dmx_ctx* ctx, long long* launches);
/* Per-kernel-class device timers inside the hot path (what Dune::Timer buckets are in newtonsolver.hh:921-955, at
   kernel granularity): CUDA-event pairs on the ctx stream around every launch of the class while enabled.
   dmx_profile(ctx, 1) resets and enables, dmx_profile(ctx, 0) disables; dmx_profile_read returns the accumulated
   device milliseconds and the number of timed units of one DMX_K_* class. */
enum { DMX_K_ASSEMBLY = 0, DMX_K_SPMV = 1, DMX_K_ILU_APPLY = 2, DMX_K_ILU_FACTOR = 3, DMX_K_AMG = 4, DMX_K_BLAS1 = 5,
       DMX_K_HALO = 6, DMX_K_JACOBI = 7 };
int  dmx_profile(dmx_ctx* ctx, int enable);
int  dmx_profile_read(dmx_ctx* ctx, int kclass, double* ms_total, long long* units);
/* device milliseconds one level of the AMG hierarchy spent in the V-cycles timed while dmx_profile was on (exclusive of the
   coarser levels) and the number of cycles */
int  dmx_amg_level_profile(dmx_ctx* ctx, int level, double* ms_total, long long* cycles);
int  dmx_synchronize(dmx_ctx* ctx);
/* device stopwatch on the ctx stream (CUDA events): start records, stop records + waits and returns the milliseconds */
int  dmx_timer_start(dmx_ctx* ctx);
int  dmx_timer_stop(dmx_ctx* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif
