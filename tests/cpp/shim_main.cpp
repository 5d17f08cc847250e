// C++ host-side drivers over include/dumux_b200.hpp, written like the reference mains:
//   mode "1p":  test/porousmediumflow/1p/incompressible/main.cc:150-161  (assembleJacobian, assembleResidual, solve, x -= dx)
//   mode "2p":  test/porousmediumflow/2p/incompressible/main.cc:126-163  (time loop, Newton with dt control)
// Prints the solution as text (one dof per line, %.17g) for the Python parity tests (tests/test_gpu_cpp_shim.py).
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

#include "dumux_b200.hpp"

using namespace dumux_b200;

static ProblemData onepIncompressible(int nx, int ny)
{
    // problem.hh:41-100, spatialparams.hh:44-106, params.input of test/porousmediumflow/1p/incompressible
    ProblemData p;
    p.model = DMX_MODEL_1P; p.dim = 2;
    p.cells = {{nx, ny, 1}};
    p.lower = {{0, 0, 0}}; p.upper = {{1, 1, 1}};
    p.permeability.assign(nx * ny, 1e-10);
    const double hx = 1.0 / nx, hy = 1.0 / ny, eps = 1.5e-7;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const double x = 0.5 * ((0.0 + i * hx) + (0.0 + (i + 1) * hx)), y = 0.5 * ((0.0 + j * hy) + (0.0 + (j + 1) * hy));
            const bool lens = !(x < 0.2 + eps || x > 0.8 - eps) && !(y < 0.2 + eps || y > 0.8 - eps);
            if (lens) p.permeability[i + nx * j] = 1e-12;
        }
    p.density[0] = 1000.0; p.viscosity[0] = 1e-3;
    // Dirichlet at y = 0 and y = 1: p = 1e5 - 1e5*(y - 1); no-flow elsewhere
    for (int s = 2; s < 4; ++s) {
        p.boundary[s].type.assign(nx, DMX_BC_DIRICHLET);
        p.boundary[s].values.assign(nx, s == 2 ? 1.0e5 + (-1.0e5) * (0.0 - 1.0) : 1.0e5 + (-1.0e5) * (1.0 - 1.0));
    }
    p.options.base_eps = 0.1;                 // Assembly.NumericDifference.BaseEpsilon
    p.options.privar_magnitude[0] = 1e5;      // Assembly.NumericDifference.PriVarMagnitude
    return p;
}

static ProblemData twopLens(int nx, int ny)
{
    // test/porousmediumflow/2p/incompressible: spatialparams.hh:46-140, problem.hh:50-168, params.input (van Genuchten)
    ProblemData p;
    p.model = DMX_MODEL_2P; p.dim = 2;
    p.cells = {{nx, ny, 1}};
    p.lower = {{0, 0, 0}}; p.upper = {{6, 4, 1}};
    p.permeability.assign(nx * ny, 4.6e-10);
    p.region.assign(nx * ny, 0);
    const double hx = 6.0 / nx, hy = 4.0 / ny, eps = 1.5e-7;
    auto xc = [&](int i) { return 0.5 * ((0.0 + i * hx) + (0.0 + (i + 1) * hx)); };
    auto yc = [&](int j) { return 0.5 * ((0.0 + j * hy) + (0.0 + (j + 1) * hy)); };
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const bool lens = !(xc(i) < 1.0 + eps || xc(i) > 4.0 - eps) && !(yc(j) < 2.0 + eps || yc(j) > 3.0 - eps);
            if (lens) { p.permeability[i + nx * j] = 9.05e-12; p.region[i + nx * j] = 1; }
        }
    p.materials.push_back({DMX_LAW_VANGENUCHTEN, {0.0037, 4.7, 0.5}, 0.05, 0.0, true, {0.01, 0.99, 0.1, 0.9}});
    p.materials.push_back({DMX_LAW_VANGENUCHTEN, {0.00045, 7.3, 0.5}, 0.18, 0.0, true, {0.01, 0.99, 0.1, 0.9}});
    const double rhoW = 1000.0, g = -9.81, height = 4.0, width = 6.0, alpha = 1 + 1.5 / height;
    auto pw = [&](double x, double y) {
        const double depth = 4.0 - y;
        const double factor = (width * alpha + (1.0 - alpha) * x) / width;
        return 1e5 - factor * rhoW * g * depth;
    };
    for (int s = 0; s < 2; ++s) {          // left / right: Dirichlet, hydrostatic-like p_w, S_n = 0
        p.boundary[s].type.assign(ny, DMX_BC_DIRICHLET);
        p.boundary[s].values.assign(2 * ny, 0.0);
        for (int j = 0; j < ny; ++j) p.boundary[s].values[2 * j] = pw(s == 0 ? 0.0 : 6.0, yc(j));
    }
    p.boundary[3].type.assign(nx, DMX_BC_NEUMANN);      // top: TCE infiltration on 2 < x < 3
    p.boundary[3].values.assign(2 * nx, 0.0);
    for (int i = 0; i < nx; ++i) {
        const double lam = (6.0 - xc(i)) / width;
        if (0.5 < lam && lam < 2.0 / 3.0) p.boundary[3].values[2 * i + 1] = -0.04;
    }
    return p;
}

int main(int argc, char** argv)
{
    const std::string mode = argc > 1 ? argv[1] : "1p";
    try {
        auto ctx = std::make_shared<Context>(0);
        if (mode == "1p-newton-ssor" || mode == "1p-newton-ilu") {
            // stationary Newton solve through GpuNewtonSolver with the solver object the caller chose: the preconditioner,
            // Krylov method, reduction and iteration limit of `linearSolver` must be the ones the device loop runs
            const int nx = 10, ny = 10;
            auto assembler = std::make_shared<GpuFVAssembler>(ctx, onepIncompressible(nx, ny));
            std::shared_ptr<GpuILUBiCGSTABSolver> linearSolver;
            if (mode == "1p-newton-ssor") linearSolver = std::make_shared<GpuSSORBiCGSTABSolver>(ctx);
            else linearSolver = std::make_shared<GpuILUBiCGSTABSolver>(ctx);
            auto other = std::make_shared<GpuILURestartedGMResSolver>(ctx);      // a second solver object on the same context must not leak into the solve
            (void)other;
            GpuNewtonSolver newton(assembler, linearSolver);
            GpuFVAssembler::SolutionVector x(assembler->numDofs(), 1, 0.0);
            newton.solve(x);
            std::fprintf(stderr, "newton %d linear %d %d\n", newton.report().newton_iterations, newton.report().linear_iterations[0],
                         newton.report().linear_iterations[1]);
            for (std::size_t i = 0; i < x.size(); ++i) std::printf("%.17g\n", x[i][0]);
        } else if (mode == "1p" || mode == "1p-ssorcg") {
            const int nx = 10, ny = 10;
            auto assembler = std::make_shared<GpuFVAssembler>(ctx, onepIncompressible(nx, ny));
            // test/porousmediumflow/1p/incompressible/main.cc uses SSORCGIstlSolver; the ILU-BiCGSTAB variant is the bench solver
            std::shared_ptr<GpuILUBiCGSTABSolver> linearSolver;
            if (mode == "1p-ssorcg") linearSolver = std::make_shared<GpuSSORCGSolver>(ctx);
            else linearSolver = std::make_shared<GpuILUBiCGSTABSolver>(ctx);
            std::fprintf(stderr, "linear solver: %s\n", linearSolver->name().c_str());
            GpuFVAssembler::SolutionVector x(assembler->numDofs(), 1, 0.0);
            assembler->setLinearSystem();
            assembler->assembleJacobian(x);
            assembler->assembleResidual(x);
            auto deltaX = x;
            auto& jacobian = assembler->jacobian();
            auto& residual = assembler->residual();
            const auto result = linearSolver->solve(jacobian, deltaX, residual);     // host-buffer form, like main.cc:156
            if (!result) throw NumericalProblem("linear solver did not converge");
            x -= deltaX;
            assembler->assembleResidual(x);
            std::fprintf(stderr, "iterations %d reduction %.3e final residual norm %.3e\n", result.iterations, result.reduction,
                         linearSolver->norm(assembler->residual()));
            for (std::size_t i = 0; i < x.size(); ++i) std::printf("%.17g\n", x[i][0]);
        } else {
            const int nx = 48, ny = 32;
            const double tEnd = 3000.0, dtInitial = 250.0;
            ProblemData pd = twopLens(nx, ny);
            GpuFVAssembler::SolutionVector x(static_cast<std::size_t>(nx) * ny, 2, 0.0);
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const double y = 0.5 * ((0.0 + j * (4.0 / ny)) + (0.0 + (j + 1) * (4.0 / ny)));
                    x[i + nx * j][0] = 1e5 - 1000.0 * (-9.81) * (4.0 - y);
                }
            auto xOld = x;
            auto assembler = std::make_shared<GpuFVAssembler>(ctx, pd, dtInitial, xOld);
            // test/porousmediumflow/2p/incompressible/main.cc:134 uses ILURestartedGMResIstlSolver
            std::shared_ptr<GpuILUBiCGSTABSolver> linearSolver;
            if (mode == "2p-gmres") linearSolver = std::make_shared<GpuILURestartedGMResSolver>(ctx);
            else linearSolver = std::make_shared<GpuILUBiCGSTABSolver>(ctx);
            std::fprintf(stderr, "linear solver: %s\n", linearSolver->name().c_str());
            GpuNewtonSolver nonLinearSolver(assembler, linearSolver);
            // plain TimeLoop (common/timeloop.hh:239-252,320-332,385-411)
            double time = 0.0, dt = dtInitial;
            auto finished = [&] { return (tEnd - time) < 1e-10 * (time - 0.0); };
            auto maxDt = [&] { return finished() ? 0.0 : std::fmax(0.0, tEnd - time); };
            dt = std::fmin(dt, maxDt());
            int steps = 0;
            do {
                dt = nonLinearSolver.solve(x, dt);
                xOld = x;
                assembler->advanceTimeStep();
                time += dt;
                ++steps;
                std::fprintf(stderr, "step %d t %.6g dt %.6g newton %d\n", steps, time, dt, nonLinearSolver.report().newton_iterations);
                dt = std::fmin(dt, maxDt());
                dt = std::fmin(nonLinearSolver.suggestTimeStepSize(dt), maxDt());
            } while (!finished());
            for (std::size_t i = 0; i < x.size(); ++i) std::printf("%.17g %.17g\n", x[i][0], x[i][1]);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
