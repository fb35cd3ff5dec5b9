// TEST INFRASTRUCTURE ONLY -- cell_population checker on top of the plain-C CVODE restatement (oracle/cvode_bdf.c)
// with the difference-quotient Jacobian of ODESolverCVODE.cpp:496-537 (jac == NULL in bdf_create).
#include <cstdlib>
#include <cstring>

#include "cellpop_glue.hpp"

extern "C" {
#include "cvode_bdf.h"
double oracle_ndtri(double p);
}

namespace cellpop_glue {
double ndtri(double p) { return oracle_ndtri(p); }

// The matching of observed to simulated cells. The reference's Hungarian implementation (dependencies/hungarian2/hungarian.cpp)
// does not return the optimal matching as the reference calls it (an integer test in its initialisation, see
// bcm3_b200/csrc/matching_host.cuh), so "any optimal matching" is not a checker for it. oracle/_ref links the reference's own
// compiled implementation and is the anchor (golden fixtures, tests/test_cellpop_cpu.py compares the restatement with it on
// random matrices); this port build has no access to the reference and takes the product's restatement -- for this one step
// the port checks the cost matrix and the plumbing around the matching, not the matching.
}
#include "../bcm3_b200/csrc/matching_host.cuh"
namespace cellpop_glue {
std::vector<int> hungarian_match(int n, int n_right, int n_left, const std::vector<double>& cost)
{
	// the reference's implementation keeps the edges whose right node index is below n (hungarian.cpp:81): with more right than
	// left nodes that is the complete graph on the first n right nodes
	if (n != n_left || n_right < n_left) return std::vector<int>();
	if (n_right == n_left) return bcm3b200::payor_matching_complete(n, cost.data());
	std::vector<double> square((size_t)n * n);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < n; j++) square[(size_t)i * n + j] = cost[(size_t)i * n_right + j];
	return bcm3b200::payor_matching_complete(n, square.data());
}
}

namespace {

struct PortSolver {
	const oracle_cellpop_problem& pr;
	bdf_mem* m;
	const double* cell_params = nullptr;
	double creation_time = 0.0;
	std::vector<double> constant_species_y;
	int64_t cnt[ORACLE_NUM_COUNTERS] = { 0 };
	void counters(int64_t* out)
	{
		for (int k = 0; k < ORACLE_NUM_COUNTERS; k++) out[k] = cnt[k];
	}
	void accumulate() // bdf_reinit zeroes the counters, as CVodeReInit does
	{
		cnt[ORACLE_CNT_NFE] += m->nfe;
		cnt[ORACLE_CNT_NSETUPS] += m->nsetups;
		cnt[ORACLE_CNT_NJE] += m->nje;
		cnt[ORACLE_CNT_NETF] += m->netf;
		cnt[ORACLE_CNT_NCFN] += m->ncfn;
		cnt[ORACLE_CNT_NNI] += m->nni;
	}
	explicit PortSolver(const oracle_cellpop_problem& p)
	    : pr(p), m(new bdf_mem), constant_species_y(p.constant_species, p.constant_species + p.num_constant_species)
	{
	}
	~PortSolver() { delete m; }
	static int rhs(double t, const double* y, double* ydot, void* user)
	{
		PortSolver* s = (PortSolver*)user;
		// Cell::SetTreatmentConcentration + solver_rhs_fn, Cell.cpp:415-433
		if (s->pr.treatment_species >= 0) s->constant_species_y[s->pr.treatment_species] = cellpop_glue::pulse_concentration(s->pr, t, s->creation_time);
		s->pr.derivative(ydot, y, s->constant_species_y.data(), s->cell_params, s->pr.non_sampled);
		return 0;
	}
	// ODESolver::SolveReturnSolution (ODESolver.cpp:93-134) + ODESolverCVODE::Solve (ODESolverCVODE.cpp:322-463) with the
	// discontinuities of a pulsed treatment (Cell.cpp:212-229, 444-460)
	bool solve(const double* y0, const double* params, const double* tp, int ntp, double* out, int& steps, double cell_creation_time,
	           cellpop_glue::SolveEvents* events)
	{
		const int N = pr.num_species;
		cell_params = params;
		creation_time = cell_creation_time;
		double next_disc = cellpop_glue::first_discontinuity_ahead(pr, creation_time);
		steps = 0;
		for (int k = 0; k < ORACLE_NUM_COUNTERS; k++) cnt[k] = 0;
		int ti = 0;
		while (tp[ti] < 2.220446049250313e-16) {
			for (int i = 0; i < N; i++) out[i + (size_t)ti * N] = y0[i];
			ti++;
			if (ti == ntp) return true;
		}
		const double end_time = tp[ntp - 1];
		bdf_create(m, N, &PortSolver::rhs, nullptr, this); // Cell::AllocateSolver happens once per cell object; state never leaks because
		                                                    // CVodeReInit resets everything the first step reads
		double atol[BDF_NMAX];
		for (int i = 0; i < N; i++) atol[i] = pr.abs_tol;
		bdf_set_tolerances(m, pr.rel_tol, atol);
		m->hmin = pr.min_dt; // SetSolverParameter("min_dt"), Cell.cpp:72
		m->hmax_inv = (pr.max_dt > 0.0 && std::isfinite(pr.max_dt)) ? 1.0 / pr.max_dt : 0.0; // "max_dt" -> CVodeSetMaxStep, cvode_io.c:344-376
		double y[BDF_NMAX], tmp[BDF_NMAX];
		for (int i = 0; i < N; i++) y[i] = y0[i];
		bdf_reinit(m, 0.0, y);
		if (!std::isnan(next_disc)) bdf_set_stop_time(m, next_disc);
		int tpi = ti;
		for (;;) {
			double tret;
			int result = bdf_step(m, end_time, y, &tret);
			if (result < 0) return false;
			steps++;
			while (tret >= tp[tpi]) {
				if (bdf_get_dky(m, tp[tpi], tmp) != BDF_SUCCESS) return false;
				for (int i = 0; i < N; i++) out[i + (size_t)tpi * N] = tmp[i];
				tpi++;
				if (tpi >= ntp) break;
			}
			if (events->after_step(tret, y, N)) break; // integration_step_cb, ODESolverCVODE.cpp:431-436
			if (tret >= end_time) break;
			if (steps == pr.max_steps) return false;
			if (!std::isnan(next_disc) && (result == BDF_TSTOP_RETURN || next_disc == tret)) { // ODESolverCVODE.cpp:448-461
				next_disc = cellpop_glue::pulse_next_discontinuity(pr, tret, creation_time);
				accumulate();
				bdf_reinit(m, tret, y);
				if (!std::isnan(next_disc) && next_disc < INFINITY) bdf_set_stop_time(m, next_disc);
			}
		}
		accumulate();
		if (const char* rep = getenv("BCM3B200_CELLPOP_REPORT")) { // debugging aid: report another counter in place of the steps
			int k = atoi(rep);
			if (k == 1) steps = (int)m->nfe; else if (k == 2) steps = (int)m->nsetups; else if (k == 3) steps = (int)m->nje;
		}
		return true;
	}
};

} // namespace

extern "C" int oracle_cellpop_evaluate(const oracle_cellpop_problem* prob, size_t num_chains, const double* values, double* logp,
                                       double* cell_values, int32_t* cell_steps, double* population_average, int num_threads)
{
	if (prob && prob->num_species > BDF_NMAX) return -3;
	return cellpop_glue::evaluate<PortSolver>(prob, num_chains, values, logp, cell_values, cell_steps, population_average, num_threads);
}

extern "C" int oracle_cellpop_evaluate_counters(const oracle_cellpop_problem* prob, size_t num_chains, const double* values, double* logp,
                                                int64_t* counters, int num_threads)
{
	if (prob && prob->num_species > BDF_NMAX) return -3;
	return cellpop_glue::evaluate<PortSolver>(prob, num_chains, values, logp, nullptr, nullptr, nullptr, num_threads, counters);
}

extern "C" int oracle_hungarian_match(int n, const double* cost, int32_t* match)
{
	std::vector<double> c(cost, cost + (size_t)n * n);
	const std::vector<int> m = cellpop_glue::hungarian_match(n, n, n, c);
	for (int i = 0; i < n; i++) match[i] = ((int)m.size() == n) ? m[i] : -1;
	return 0;
}
