// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- cell_population checker around the reference's REAL compiled
// ODESolverCVODE (difference-quotient Jacobian, zero-skipping LU, CVODE 5.3.0); the glue is oracle/cellpop_glue.hpp.
#include "Utils.h"
#include "ODESolverCVODE.h"

#include "../cellpop_glue.hpp"

namespace refglue {
double ndtri(double p); // poppk_ref.cpp
}
namespace cellpop_glue {
double ndtri(double p) { return refglue::ndtri(p); }
}

namespace {

struct RefSolver {
	const oracle_cellpop_problem& pr;
	ODESolverCVODE solver;
	const double* cell_params = nullptr;
	explicit RefSolver(const oracle_cellpop_problem& p) : pr(p)
	{
		// Cell::AllocateSolver, Cell.cpp:57-76 (no SetJacobianFunction => DifferenceQuotientJacobian)
		solver.SetDerivativeFunction([this](OdeReal, const OdeReal* y, OdeReal* ydot, void*) {
			pr.derivative(ydot, y, pr.constant_species, cell_params, pr.non_sampled);
			return true;
		});
		solver.Initialize((size_t)pr.num_species, nullptr, 0);
		solver.SetTolerance(pr.rel_tol, pr.abs_tol);
		solver.SetSolverParameter("min_dt", 0, pr.min_dt);
		solver.SetSolverParameter("max_dt", 0, std::numeric_limits<OdeReal>::infinity());
		solver.SetSolverParameter("max_steps", pr.max_steps, std::numeric_limits<OdeReal>::quiet_NaN());
	}
	bool solve(const double* y0, const double* params, const double* tp, int ntp, double* out, int& steps)
	{
		cell_params = params;
		Eigen::Map<const OdeVectorReal> ic(y0, pr.num_species);
		OdeVectorReal initial = ic;
		OdeVectorReal timepoints = Eigen::Map<const OdeVectorReal>(tp, ntp);
		OdeMatrixReal output;
		bool ok = solver.SolveReturnSolution(initial, &timepoints, &output);
		steps = (int)solver.GetNumSteps();
		if (ok) {
			for (int t = 0; t < ntp; t++)
				for (int i = 0; i < pr.num_species; i++) out[i + (size_t)t * pr.num_species] = output(i, t);
		}
		return ok;
	}
};

} // namespace

extern "C" int oracle_cellpop_evaluate(const oracle_cellpop_problem* prob, size_t num_chains, const double* values, double* logp,
                                       double* cell_values, int32_t* cell_steps, double* population_average, int num_threads)
{
	return cellpop_glue::evaluate<RefSolver>(prob, num_chains, values, logp, cell_values, cell_steps, population_average, num_threads);
}
