// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- cell_population checker around the reference's REAL compiled
// ODESolverCVODE (difference-quotient Jacobian, zero-skipping LU, CVODE 5.3.0); the glue is oracle/cellpop_glue.hpp.
#include "Utils.h"
#include "ODESolverCVODE.h"

#include "../cellpop_glue.hpp"

namespace refglue {
double ndtri(double p); // poppk_ref.cpp
}
#include "hungarian.h" // the reference's dependencies/hungarian2 (compiled in place by the Makefile)

namespace cellpop_glue {
double ndtri(double p) { return refglue::ndtri(p); }
// the edge list exactly as DataLikelihoodTimeCourse::Evaluate builds it (.cpp:288-323): every (observed, simulated) pair in row order
std::vector<int> hungarian_match(int n, int n_right, int n_left, const std::vector<double>& cost)
{
	std::vector<WeightedBipartiteEdge> edges((size_t)n_left * n_right);
	int edge_count = 0;
	for (int i = 0; i < n_left; i++)
		for (int j = 0; j < n_right; j++) {
			edges[edge_count].left = i;
			edges[edge_count].right = j;
			edges[edge_count].cost = cost[(size_t)i * n_right + j];
			edge_count++;
		}
	return hungarianMinimumWeightPerfectMatching(n, n_right, edges, edge_count);
}
}

namespace {

// subclass only to reach the protected cvode_mem for the statistics counters (CVodeReInit zeroes them: they are read before
// every re-initialisation and at the end of the solve)
struct CountingCVODE : public ODESolverCVODE {
	void Accumulate(int64_t* cnt)
	{
		long v;
		if (CVodeGetNumRhsEvals(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NFE] += v;
		if (CVodeGetNumLinSolvSetups(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NSETUPS] += v;
		if (CVodeGetNumErrTestFails(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NETF] += v;
		if (CVodeGetNumNonlinSolvConvFails(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NCFN] += v;
		if (CVodeGetNumNonlinSolvIters(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NNI] += v;
		if (CVodeGetNumJacEvals(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NJE] += v;
	}
};

struct RefSolver {
	const oracle_cellpop_problem& pr;
	CountingCVODE solver;
	int64_t cnt[ORACLE_NUM_COUNTERS] = { 0 };
	void counters(int64_t* out)
	{
		for (int k = 0; k < ORACLE_NUM_COUNTERS; k++) out[k] = cnt[k];
	}
	const double* cell_params = nullptr;
	double creation_time = 0.0;
	std::vector<double> constant_species_y;
	explicit RefSolver(const oracle_cellpop_problem& p) : pr(p), constant_species_y(p.constant_species, p.constant_species + p.num_constant_species)
	{
		// Cell::AllocateSolver, Cell.cpp:57-76 (no SetJacobianFunction => DifferenceQuotientJacobian)
		solver.SetDerivativeFunction([this](OdeReal t, const OdeReal* y, OdeReal* ydot, void*) {
			// Cell::SetTreatmentConcentration + solver_rhs_fn, Cell.cpp:415-433
			if (pr.treatment_species >= 0) constant_species_y[pr.treatment_species] = cellpop_glue::pulse_concentration(pr, t, creation_time);
			pr.derivative(ydot, y, constant_species_y.data(), cell_params, pr.non_sampled);
			return true;
		});
		solver.Initialize((size_t)pr.num_species, nullptr, 0);
		solver.SetTolerance(pr.rel_tol, pr.abs_tol);
		solver.SetSolverParameter("min_dt", 0, pr.min_dt);
		solver.SetSolverParameter("max_dt", 0, pr.max_dt); // Cell.cpp:73 (infinity unless the experiment sets solver_max_timestep)
		solver.SetSolverParameter("max_steps", pr.max_steps, std::numeric_limits<OdeReal>::quiet_NaN());
	}
	bool solve(const double* y0, const double* params, const double* tp, int ntp, double* out, int& steps, double cell_creation_time,
	           cellpop_glue::SolveEvents* events)
	{
		cell_params = params;
		creation_time = cell_creation_time;
		solver.Restart(); // Cell::Initialize, Cell.cpp:187
		// Cell::AllocateSolver installs the callback once (Cell.cpp:64-66); here it is re-pointed at this cell's record
		ODESolver::TIntegrationStepCallback step_cb = [this, events](OdeReal t, const OdeReal* y, Real& end_time, void*) -> bool {
			return !events->after_step(t, y, pr.num_species);
		};
		solver.SetIntegrationStepCallback(step_cb);
		for (int k = 0; k < ORACLE_NUM_COUNTERS; k++) cnt[k] = 0;
		// Cell::Simulate, Cell.cpp:212-229 + discontinuity_cb :444-460
		const double first = cellpop_glue::first_discontinuity_ahead(pr, creation_time);
		if (!std::isnan(first)) {
			ODESolver::TDiscontinuityCallback cb = [this](OdeReal t, void*) -> Real {
				solver.Accumulate(cnt); // the callback runs right before CVodeReInit (ODESolverCVODE.cpp:448-457)
				return cellpop_glue::pulse_next_discontinuity(pr, t, creation_time);
			};
			solver.SetDiscontinuity(first, cb, nullptr);
		}
		Eigen::Map<const OdeVectorReal> ic(y0, pr.num_species);
		OdeVectorReal initial = ic;
		OdeVectorReal timepoints = Eigen::Map<const OdeVectorReal>(tp, ntp);
		OdeMatrixReal output;
		bool ok = solver.SolveReturnSolution(initial, &timepoints, &output);
		steps = (int)solver.GetNumSteps();
		solver.Accumulate(cnt);
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

extern "C" int oracle_cellpop_evaluate_counters(const oracle_cellpop_problem* prob, size_t num_chains, const double* values, double* logp,
                                                int64_t* counters, int num_threads)
{
	return cellpop_glue::evaluate<RefSolver>(prob, num_chains, values, logp, nullptr, nullptr, nullptr, num_threads, counters);
}

/* the reference's compiled hungarianMinimumWeightPerfectMatching on a complete [n][n] cost matrix, edges in row order as
 * DataLikelihoodTimeCourse::Evaluate builds them; match [n] (-1 everywhere when it returns no matching). */
extern "C" int oracle_hungarian_match(int n, const double* cost, int32_t* match)
{
	std::vector<double> c(cost, cost + (size_t)n * n);
	const std::vector<int> m = cellpop_glue::hungarian_match(n, n, n, c);
	for (int i = 0; i < n; i++) match[i] = ((int)m.size() == n) ? m[i] : -1;
	return 0;
}
