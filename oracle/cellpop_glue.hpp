// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Restatement of the Boost/NetCDF-bound glue of the cell_population likelihood around an ODE solver adapter:
//   CellPopulationLikelihood::EvaluateLogProbability   src/cellpop/CellPopulationLikelihood.cpp:82-101
//   Experiment::EvaluateLogProbability / Simulate      src/cellpop/Experiment.cpp:239-372, 635-724
//   CellPopulation::AddNewCell                         src/cellpop/CellPopulation.cpp:36-104 (Sobol row = cell index)
//   Cell::Initialize / Simulate                        src/cellpop/Cell.cpp:150-273 (no events, no treatment trajectories)
//   VariabilityDescription::GetPseudorandomVector      src/cellpop/VariabilityDescription.cpp:50-64 (diagonal_gaussian)
//   VariabilityDescriptionVariable::Apply*             src/cellpop/VariabilityDescriptionVariable.cpp:80-110,172-207
//   DataLikelihoodTimeCoursePopulationAverage          .cpp:85-197;  DataLikelihoodTimeCourseBase .cpp:229-315
// Used twice: oracle/ref/cellpop_ref.cpp plugs in the reference's real ODESolverCVODE, oracle/cellpop_port.cpp plugs in
// oracle/cvode_bdf.c. The adapter must provide
//   bool solve(const double* y0, const double* cell_params, const double* timepoints_rel, int ntp, double* out /*[N][ntp] col-major*/, int& steps,
//              double creation_time, SolveEvents* events)
// with the semantics of ODESolver::SolveReturnSolution and, after every accepted step, of Cell::integration_step_cb
// (events->after_step: true = stop). Dividing cells: Experiment.cpp:726-782, CellPopulation.cpp:36-104, Cell.cpp:119-148.
#pragma once

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

#include "oracle_api.h"

namespace cellpop_glue {

double ndtri(double p); // defined by the including translation unit
// minimum-cost perfect matching of n_left left nodes (rows of cost [n_left][n_right]) to right nodes, with the call shape of the
// reference's hungarianMinimumWeightPerfectMatching(n, n_right, edges, edge_count): defined by the including translation unit
// (oracle/ref: the reference's own dependencies/hungarian2/hungarian.cpp; oracle/cellpop_port.cpp: a restatement)
std::vector<int> hungarian_match(int n, int n_right, int n_left, const std::vector<double>& cost);

inline double transform_variable(int tr, double x)
{
	switch (tr) {
	case ORACLE_TRANSFORM_LOG: return exp(x);
	case ORACLE_TRANSFORM_LOG10: return exp(x * 2.3025850929940459);
	case ORACLE_TRANSFORM_LOGIT:
		if (x > 0) { double z = exp(-x); return 1.0 / (1.0 + z); } else { double z = exp(x); return z / (1.0 + z); }
	default: return x;
	}
}

inline void apply_variability(double& x, double value, int apply)
{
	switch (apply) {
	case 0: x += value; break;
	case 1: x += exp(value); break;
	case 2: x += pow(2.0, value); break;
	case 3: x *= value; break;
	case 4: x *= exp(value); break;
	case 5: x *= pow(2.0, value); break;
	case 6: x = value; break;
	}
}

inline double logpdf_tnu4(double x, double mu, double sigma)
{
	double xn = (x - mu) / sigma;
	return -0.9808292530117262 - 2.5 * log1p(0.25 * xn * xn) - log(sigma);
}
inline double logpdf_normal(double x, double mu, double sigma)
{
	double d = x - mu;
	return -log(sigma) - 0.91893853320467274178032973640562 - d * d / (2.0 * sigma * sigma);
}

// TreatmentTrajectoryPulses::{GetConcentration, FirstDiscontinuity, NextDiscontinuity} (TreatmentTrajectoryPulses.cpp:21-71):
// each pulse starts 2 h after its time point: linear rise over 2 h, plateau 8 h, linear decay over 4 h.
inline double pulse_concentration(const oracle_cellpop_problem& pr, double time, double cell_creation_time)
{
	const double global_time = time + cell_creation_time;
	for (int i = 0; i < pr.treatment_num_pulses; i++) {
		const double t_in_pulse = global_time - pr.treatment_times[i] - 2.0;
		if (t_in_pulse >= 14.0) continue;
		else if (t_in_pulse <= 0.0) return 0.0;
		else if (t_in_pulse < 2.0) return t_in_pulse * 0.5;
		else if (t_in_pulse < 10.0) return 1.0;
		else return 1 - (t_in_pulse - 10.0) * 0.25;
	}
	return 0.0;
}
inline double pulse_first_discontinuity(const oracle_cellpop_problem& pr, double cell_creation_time)
{
	if (pr.treatment_num_pulses > 0) return pr.treatment_times[0] - cell_creation_time + 2.0;
	return std::numeric_limits<double>::quiet_NaN();
}
inline double pulse_next_discontinuity(const oracle_cellpop_problem& pr, double time, double cell_creation_time)
{
	for (int i = 0; i < pr.treatment_num_pulses; i++) {
		const double* tp = pr.treatment_times;
		if (time == tp[i] - cell_creation_time + 2.0) return tp[i] - cell_creation_time + 4.0;
		else if (time == tp[i] - cell_creation_time + 4.0) return tp[i] - cell_creation_time + 10.0;
		else if (time == tp[i] - cell_creation_time + 10.0) return tp[i] - cell_creation_time + 14.0;
		else if (time == tp[i] - cell_creation_time + 14.0) {
			if (i < pr.treatment_num_pulses - 1) return tp[i + 1] - cell_creation_time + 2.0;
			else return std::numeric_limits<double>::quiet_NaN();
		}
	}
	return std::numeric_limits<double>::quiet_NaN();
}
// Cell::Simulate, Cell.cpp:212-229: the first discontinuity that lies ahead of the cell (NaN: none)
inline double first_discontinuity_ahead(const oracle_cellpop_problem& pr, double cell_creation_time)
{
	if (pr.treatment_species < 0) return std::numeric_limits<double>::quiet_NaN();
	double discontinuity = pulse_first_discontinuity(pr, cell_creation_time);
	if (!std::isnan(discontinuity)) {
		while (discontinuity < 0.0) discontinuity = pulse_next_discontinuity(pr, discontinuity, cell_creation_time);
	}
	if (!std::isnan(discontinuity) && discontinuity > 0.0) return discontinuity;
	return std::numeric_limits<double>::quiet_NaN();
}

// What the integration-step callback of a cell reports (Cell::integration_step_cb, Cell.cpp:463-538, the branch without stored
// integration points): the integration ended at the step that took "cytokinesis" / "apoptosis" above 1, at cell time t_event
// with state y_event (the solver's y at the end of that step).
struct SolveEvents {
	int cytokinesis_ix = -1, apoptosis_ix = -1;
	// Cell.cpp:487-492: the first accepted step after which the nuclear envelope species is below 0.5 sets the breakdown time (to
	// whatever get_threshold_crossing_time returns: always a number), and with it Cell::EnteredMitosis (Cell.h:27)
	int nuclear_envelope_ix = -1;
	bool entered_mitosis = false;
	bool divided = false, died = false;
	double t_event = 0.0;
	std::vector<double> y_event;
	// true = stop integrating
	bool after_step(double t, const double* y, int N)
	{
		bool stop = false;
		if (nuclear_envelope_ix >= 0 && y[nuclear_envelope_ix] < 0.5) entered_mitosis = true;
		if (cytokinesis_ix >= 0 && y[cytokinesis_ix] > 1.0) {
			divided = true;
			stop = true;
		}
		if (apoptosis_ix >= 0 && y[apoptosis_ix] > 1.0) {
			died = true;
			stop = true;
		}
		if (stop) {
			t_event = t;
			y_event.assign(y, y + N);
		}
		return stop;
	}
};

template <class Solver>
void evaluate_chain(const oracle_cellpop_problem& pr, const double* values, double* logp_out, double* cell_values, int32_t* cell_steps,
                    double* pop_avg_out, int64_t* cell_counters = nullptr, int cell_threads = 1)
{
	const int N = pr.num_species, nvar = pr.num_variables, T = pr.num_timepoints, ninit = pr.num_cells, D = pr.variability_dim, R = pr.num_replicates;
	const bool dividing = pr.divide_cells != 0;
	const int ncell = dividing ? pr.max_cells : ninit; // capacity = number of cell columns of the outputs
	const double nan = std::numeric_limits<double>::quiet_NaN();
	std::vector<double> transformed(nvar);
	for (int i = 0; i < nvar; i++) transformed[i] = transform_variable(pr.transforms[i], values[i]);
	const double entry_time = (pr.entry_time_ix >= 0) ? transformed[pr.entry_time_ix] : pr.entry_time;

	// VariabilityDescription::GetPseudorandomVector, FullGaussian (VariabilityDescription.cpp:99-131): spherical
	// parametrisation of the Cholesky factor, L(i, j) = exp(scale_i) * prod_{k < j} sin(pi c_ik) * [j < i] cos(pi c_ij)
	std::vector<double> cholesky((size_t)D * D, 0.0);
	if (pr.full_gaussian) {
		for (int i = 0; i < D; i++) {
			const double* row = pr.variability + (size_t)i * 6;
			const int scale_ix = (int)row[3];
			const double exp_scale = exp((scale_ix >= 0) ? transformed[scale_ix] : row[4]);
			for (int j = 0; j <= i; j++) {
				double l = exp_scale;
				for (int k = 0; k < i; k++) {
					if (k <= j) {
						const double* cv = pr.covariance + (size_t)((i - 1) * i / 2 + k) * 2;
						const double cov_value = (((int)cv[0] >= 0) ? transformed[(int)cv[0]] : cv[1]) * M_PI;
						if (k == j) l *= cos(cov_value);
						else l *= sin(cov_value);
					}
				}
				cholesky[(size_t)i * D + j] = l;
			}
		}
	}

	// the solver integrates to the last output time it is given: when the experiment runs longer than this data set, one
	// extra output time (whose values nobody reads) carries the experiment's end
	const bool longer = pr.have_sim_end_time && pr.sim_end_time > pr.timepoints[T - 1];
	const int Tsolve = T + (longer ? 1 : 0);
	const double target_time = longer ? pr.sim_end_time : pr.timepoints[T - 1]; // Experiment::Simulate's simulation_end_time
	std::vector<double> population_average(T, 0.0);
	std::vector<double> xs((size_t)T * ncell, nan); // value per (timepoint, cell)
	const int L = 1 + ((pr.data_kind != 0) ? pr.num_extra_markers : 0); // markers of a per-cell data set
	std::vector<std::vector<double>> xs_marker((size_t)(L - 1), std::vector<double>((size_t)T * ncell, nan));

	// The population (CellPopulation): cells in creation order. Initial cells take the quasi-random rows 0 .. ninit - 1, the
	// daughters of the cell with row r take ninit + 2 r + child (CellPopulation.cpp:56-80).
	struct CellRec {
		double creation = 0.0, sim_end = 0.0; // sim_end: cell time up to which the cell exists (Cell::simulation_end_time)
		long row = 0;
		int parent = -1;
		bool initial = false, ok = true, divided = false, died = false, entered_mitosis = false;
		std::vector<double> y0, end_y;
		double achieved = 0.0;
	};
	std::vector<CellRec> cells;
	cells.reserve(ncell);
	for (int ci = 0; ci < ninit; ci++) {
		CellRec c;
		c.creation = entry_time;
		c.row = ci;
		c.initial = ninit > 1; // Experiment.cpp:662-670: the flag handed to Cell::Initialize is false for a single initial cell
		c.y0.assign(pr.initial_conditions, pr.initial_conditions + N);
		cells.push_back(std::move(c));
	}
	bool result = true;

	// Cell::Initialize + Cell::Simulate of one cell; everything it writes is its own (cell record, its column of xs, its counters)
	auto simulate_cell = [&](Solver& solver, int ci) {
		CellRec& c = cells[ci];
		std::vector<double> cell_params(transformed), y0(c.y0), tp_rel(Tsolve), out((size_t)N * Tsolve);
		for (int d = 0; d < D; d++) {
			const double* row = pr.variability + (size_t)d * 6;
			const int kind = (int)row[0] & 3;
			const bool only_initial = ((int)row[0] & 4) != 0; // VariabilityDescriptionVariable.cpp:66-110
			const bool is_ic = kind == 1;
			const int target = (int)row[1], apply = (int)row[2], scale_ix = (int)row[3];
			double v;
			if (pr.full_gaussian) {
				v = 0.0; // v = L z, row d
				for (int j = 0; j <= d; j++) v += cholesky[(size_t)d * D + j] * ndtri(pr.sobol[(size_t)c.row * D + j]);
			} else {
				const double scale = (scale_ix >= 0) ? transformed[scale_ix] : row[4];
				v = ndtri(pr.sobol[(size_t)c.row * D + d]) * exp(scale);
			}
			if (row[5] != 0.0) v = -v;
			if (kind == 2) continue; // entry-time variable: takes a dimension, is never applied (no caller of ApplyVariabilityEntryTime)
			if (only_initial && !c.initial) continue;
			if (is_ic) apply_variability(y0[target], v, apply);
			else apply_variability(cell_params[target], v, apply);
		}
		// Cell::Simulate: output times relative to the creation time
		const double creation_time = c.creation;
		for (int i = 0; i < T; i++) tp_rel[i] = pr.timepoints[i] - creation_time;
		if (longer) tp_rel[T] = pr.sim_end_time - creation_time;
		int steps = 0;
		SolveEvents ev;
		if (dividing) ev.cytokinesis_ix = pr.cytokinesis_ix;
		ev.apoptosis_ix = pr.apoptosis_ix;
		ev.nuclear_envelope_ix = pr.nuclear_envelope_ix;
		c.ok = solver.solve(y0.data(), cell_params.data(), tp_rel.data(), Tsolve, out.data(), steps, creation_time, &ev);
		c.divided = ev.divided;
		c.died = ev.died;
		c.entered_mitosis = ev.entered_mitosis;
		// Cell.cpp:243-256: a cell that divided or died exists up to the event, any other up to the last requested time
		c.sim_end = (ev.divided || ev.died) ? ev.t_event : std::max(target_time - creation_time, tp_rel[Tsolve - 1]);
		c.achieved = ((ev.divided || ev.died) ? ev.t_event : tp_rel[Tsolve - 1]) + creation_time;
		if (ev.divided) c.end_y = ev.y_event;
		if (cell_steps) cell_steps[ci] = steps;
		if (cell_counters) { // per cell: ORACLE_CNT_* (steps, nfe, nsetups, nje, netf, ncfn, nni, ok), summed over the restarts of the solve
			int64_t* k = cell_counters + (size_t)ci * ORACLE_NUM_COUNTERS;
			solver.counters(k);
			k[ORACLE_CNT_STEPS] = steps;
			k[ORACLE_CNT_OK] = c.ok ? 1 : 0;
		}
		if (c.ok) {
			// Experiment.cpp:298-312 + Cell::GetInterpolatedSpeciesValue (Cell.cpp:280-360): exact stored timepoints only
			for (int i = 0; i < T; i++) {
				const double cell_time = pr.timepoints[i] - creation_time;
				if (cell_time < 0.0 || cell_time > c.sim_end) continue;
				double x = 0.0;
				for (int k = 0; k < pr.num_obs_species; k++) x += out[(size_t)pr.obs_species[k] + (size_t)i * N];
				// use_log_ratio (NotifySimulatedValue, DataLikelihoodTimeCourse.cpp:380-397): the numerator is notified first (val = a), the
				// denominator then turns it into bcm3::log10(a / b), b replaced by 1e-16 when it is smaller (MathFunctions.h:11)
				auto log_ratio = [](double num, double den) {
					return 0.4342944819032518276511289189166 * ((den < 1e-16) ? log(num / 1e-16) : log(num / den));
				};
				if (pr.data_kind == 1 && pr.log_ratio_denominator >= 0) x = log_ratio(x, out[(size_t)pr.log_ratio_denominator + (size_t)i * N]);
				xs[(size_t)i * ncell + ci] = x;
				for (int l = 1; l < L; l++) {
					const oracle_cellpop_marker& mk = pr.extra_markers[l - 1];
					double xm = 0.0;
					for (int k = 0; k < mk.num_obs_species; k++) xm += out[(size_t)mk.obs_species[k] + (size_t)i * N];
					if (pr.data_kind == 1 && mk.log_ratio_denominator >= 0) xm = log_ratio(xm, out[(size_t)mk.log_ratio_denominator + (size_t)i * N]);
					xs_marker[(size_t)(l - 1)][(size_t)i * ncell + ci] = xm;
				}
			}
		}
	};

	// Experiment::ParallelSimulation + SimulateCell (Experiment.cpp:691-782), one generation at a time: the cells of a generation
	// are independent, and creating the daughters in the order of their parents afterwards gives the cell indices the
	// reference's single-threaded loop gives (it walks the population by index and appends daughters at the end).
	std::vector<std::unique_ptr<Solver>> solvers;
	for (int t = 0; t < std::max(1, cell_threads); t++) solvers.emplace_back(new Solver(pr));
	size_t wave_begin = 0;
	while (result && wave_begin < cells.size()) {
		const size_t wave_end = cells.size();
		if (solvers.size() == 1) {
			for (size_t ci = wave_begin; ci < wave_end; ci++) simulate_cell(*solvers[0], (int)ci);
		} else {
			std::atomic<size_t> next(wave_begin);
			std::vector<std::thread> th;
			for (size_t t = 0; t < solvers.size(); t++) {
				th.emplace_back([&, t]() {
					for (;;) {
						const size_t ci = next.fetch_add(1);
						if (ci >= wave_end) break;
						simulate_cell(*solvers[t], (int)ci);
					}
				});
			}
			for (auto& t : th) t.join();
		}
		for (size_t ci = wave_begin; ci < wave_end && result; ci++) {
			const CellRec& c = cells[ci];
			if (!c.ok) {
				result = false; // Experiment::Simulate fails => logp = -inf (Experiment.cpp:356-358)
				if (getenv("ORACLE_CELLPOP_DEBUG")) fprintf(stderr, "cell %zu (row %ld, parent %d, created %.6g) failed in the solver\n", ci, c.row, c.parent, c.creation);
				break;
			}
			if (dividing && c.divided && c.achieved < target_time) {
				for (int child = 0; child < 2; child++) {
					CellRec d;
					d.creation = c.achieved;
					d.parent = (int)ci;
					d.row = (long)ninit + c.row * 2 + child;
					d.initial = false;
					// CellPopulation::AddNewCell: no free cell object, or no quasi-random row left => the evaluation fails
					if ((int)cells.size() >= ncell || d.row >= (long)pr.sobol_rows) {
						if (getenv("ORACLE_CELLPOP_DEBUG")) fprintf(stderr, "no room for a daughter of cell %zu: %zu cells of %d, row %ld of %d\n", ci, cells.size(), ncell, d.row, pr.sobol_rows);
						result = false;
						break;
					}
					// Cell::SetInitialConditionsFromOtherCell (Cell.cpp:119-148): the parent's state, seven species reset
					d.y0 = c.end_y;
					static const double reset_value[7] = { 0.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0 };
					for (int k = 0; k < 7; k++) d.y0[pr.reset_ix[k]] = reset_value[k];
					cells.push_back(std::move(d));
				}
			}
		}
		wave_begin = wave_end;
	}
	const int nactive = (int)cells.size();
	if (cell_values)
		for (size_t e = 0; e < xs.size(); e++) cell_values[e] = xs[e];

	if (!result) {
		*logp_out = -std::numeric_limits<double>::infinity();
		if (pop_avg_out) for (int i = 0; i < T; i++) pop_avg_out[i] = nan;
		return;
	}
	// NotifySimulatedValue (.cpp:161-197): value / population size, accumulated cell by cell; CountCellsAtTime (CellPopulation.cpp:
	// 106-121, Cell::CellAliveAtTime Cell.cpp:362-396): the cells whose life span [0, simulation_end_time] covers the time
	for (int i = 0; i < T; i++) {
		const bool only_mitotic = pr.include_only_mitotic != 0; // .cpp:171-176: these cells only, over their own number
		size_t pop = 0;
		for (int ci = 0; ci < nactive; ci++) {
			const double cell_time = pr.timepoints[i] - cells[ci].creation;
			if (!(cell_time < 0.0 || cell_time > cells[ci].sim_end) && (!only_mitotic || cells[ci].entered_mitosis)) pop++;
		}
		for (int ci = 0; ci < nactive; ci++) {
			const double x = xs[(size_t)i * ncell + ci];
			if (x == x && (!only_mitotic || cells[ci].entered_mitosis)) population_average[i] += x / (double)pop; // the accumulator starts at zero (Reset, .cpp:77-83)
		}
	}
	if (pop_avg_out) for (int i = 0; i < T; i++) pop_avg_out[i] = population_average[i];

	// per-marker view of a per-cell data set: marker 0 is the problem's own entries, the others come from extra_markers
	struct MarkerView {
		const double* xs;       // [T][ncell] species sum per (timepoint, cell)
		const double* observed; // [observed cells][T]
		double stdev, offset, scale, prop_stdev;
	};
	std::vector<MarkerView> markers;
	if (pr.data_kind != 0) {
		for (int l = 0; l < L; l++) {
			MarkerView mv;
			if (l == 0) {
				mv.xs = xs.data();
				mv.observed = pr.observed;
				mv.stdev = (pr.stdev_ix >= 0) ? transformed[pr.stdev_ix] : pr.stdev;
				mv.offset = (pr.offset_ix >= 0) ? transformed[pr.offset_ix] : pr.offset;
				mv.scale = (pr.scale_ix >= 0) ? transformed[pr.scale_ix] : pr.scale;
				mv.prop_stdev = (pr.proportional_stdev_ix >= 0) ? transformed[pr.proportional_stdev_ix] : pr.proportional_stdev;
			} else {
				const oracle_cellpop_marker& mk = pr.extra_markers[l - 1];
				mv.xs = xs_marker[(size_t)(l - 1)].data();
				mv.observed = mk.observed;
				mv.stdev = (mk.stdev_ix >= 0) ? transformed[mk.stdev_ix] : mk.stdev;
				mv.offset = (mk.offset_ix >= 0) ? transformed[mk.offset_ix] : mk.offset;
				mv.scale = (mk.scale_ix >= 0) ? transformed[mk.scale_ix] : mk.scale;
				mv.prop_stdev = (mk.proportional_stdev_ix >= 0) ? transformed[mk.proportional_stdev_ix] : mk.proportional_stdev;
			}
			if (pr.stdev_relative_to_scale) mv.stdev *= mv.scale; // GetCurrentSTDev(.., i), DataLikelihoodBase.cpp:151-153
			markers.push_back(mv);
		}
	}

	if (pr.data_kind == 1) {
		// <data type="time_course">: DataLikelihoodTimeCourse::Evaluate (.cpp:230-365) with CalculateCellLikelihood (.cpp:431-505) and
		// CalculateMissingValueLikelihood (.cpp:566-588) -- synchronize="none", no parent information. The trajectory of simulated
		// cell j for marker l is column j of that marker's xs (NotifySimulatedValue .cpp:367-404 adds the species of a sum);
		// observed cell i is row i of the marker's `observed`.
		const int n_obs = R, n_sim = nactive;
		std::vector<std::vector<double>> traj((size_t)L), min_log_sigma((size_t)L), inv_two_sigma_sq_cell((size_t)L);
		for (int l = 0; l < L; l++) {
			const MarkerView& mv = markers[(size_t)l];
			traj[(size_t)l].resize((size_t)T * n_sim);
			for (int k = 0; k < T; k++)
				for (int j = 0; j < n_sim; j++) {
					double v = mv.xs[(size_t)k * ncell + j];
					v *= mv.scale; // .cpp:236-241
					v += mv.offset;
					if (pr.saturation_scale_ix >= 0) { // .cpp:243-254, operation by operation
						const double saturation_scale = transformed[pr.saturation_scale_ix];
						v *= -1.0;
						v = exp(v);
						v += 1.0;
						v = 1.0 / v;
						v *= saturation_scale;
						v -= 0.5 * saturation_scale;
					}
					traj[(size_t)l][(size_t)k * n_sim + j] = v;
				}
			if (pr.error_model == 2 || pr.error_model == 3) { // .cpp:272-283
				min_log_sigma[(size_t)l].resize(traj[(size_t)l].size());
				inv_two_sigma_sq_cell[(size_t)l].resize(traj[(size_t)l].size());
				for (size_t e = 0; e < traj[(size_t)l].size(); e++) {
					double sigma = mv.prop_stdev * std::max(traj[(size_t)l][e], 0.0);
					if (pr.error_model == 3) sigma += mv.stdev;
					min_log_sigma[(size_t)l][e] = -log(sigma);
					inv_two_sigma_sq_cell[(size_t)l][e] = 1.0 / (2.0 * (sigma * sigma));
				}
			}
		}
		auto missing_value = [&](int l, int j, int k) { // .cpp:566-588
			const std::vector<double>& tr = traj[(size_t)l];
			double first_ok = pr.timepoints[T - 1], last_ok = pr.timepoints[0];
			for (int m = 0; m < T; m++) if (!std::isnan(tr[(size_t)m * n_sim + j])) { first_ok = pr.timepoints[m]; break; }
			for (int m = T - 1; m >= 0; m--) if (!std::isnan(tr[(size_t)m * n_sim + j])) { last_ok = pr.timepoints[m]; break; }
			const double time_offset = std::min(std::abs(pr.timepoints[k] - first_ok), std::abs(pr.timepoints[k] - last_ok));
			return (pr.error_model == 1) ? logpdf_tnu4(time_offset, 0, pr.missing_stdev) : logpdf_normal(time_offset, 0, pr.missing_stdev);
		};
		const int n = std::max(n_obs, n_sim);
		std::vector<double> cell_likelihoods((size_t)n * n, 0.0);
		const double ninf = -std::numeric_limits<double>::infinity();
		for (int i = 0; i < n_obs; i++) {
			int finite_count = 0;
			for (int j = 0; j < n_sim; j++) {
				double cell_logp = 0.0;
				for (int l = 0; l < L; l++) {
					const MarkerView& mv = markers[(size_t)l];
					const std::vector<double>& tr = traj[(size_t)l];
					double opt_offset = 0.0, opt_scale = 1.0;
					if (pr.optimize_offset_scale) {
						// OptimizeOffsetScale (DataLikelihoodTimeCourseBase.cpp:317-322) = bcm3::linear_regress_columns(simulated, observed)
						// (Correlation.cpp:158-200: running means, NaN pairs skipped) + the clamps
						double mu_x = 0.0, mu_y = 0.0, xvar_calc = 0.0, cov_calc = 0.0, empirical_n = 0.0;
						for (int k = 0; k < T; k++) {
							const double xv = tr[(size_t)k * n_sim + j], yv = mv.observed[(size_t)i * T + k];
							if (std::isnan(xv) || std::isnan(yv)) continue;
							empirical_n += 1.0;
							const double invN = 1.0 / empirical_n, mu_x_nm1 = mu_x, mu_y_nm1 = mu_y;
							mu_x += (xv - mu_x) * invN;
							mu_y += (yv - mu_y) * invN;
							if (empirical_n > 1) {
								const double ratio = (empirical_n - 1) / empirical_n, dx = xv - mu_x_nm1, dy = yv - mu_y_nm1;
								xvar_calc += dx * dx * ratio;
								cov_calc += dx * dy * ratio;
							}
						}
						if (empirical_n >= 2) {
							opt_scale = cov_calc / xvar_calc;
							opt_offset = mu_y - mu_x * opt_scale;
						}
						opt_scale = std::min(std::max(opt_scale, pr.optimize_scale_min), pr.optimize_scale_max);
						opt_offset = std::min(std::max(opt_offset, pr.optimize_offset_min), pr.optimize_offset_max);
					}
					const double minus_log_sigma = -log(mv.stdev), inv_two_sigma_sq = 1.0 / (2.0 * mv.stdev * mv.stdev); // .cpp:452-453
					for (int k = 0; k < T; k++) {
						const double y = mv.observed[(size_t)i * T + k];
						if (std::isnan(y)) continue;
						const double x = opt_offset + opt_scale * tr[(size_t)k * n_sim + j]; // .cpp:461
						if (std::isnan(x)) {
							cell_logp += missing_value(l, j, k);
						} else if (pr.error_model == 0) {
							const double d = y - x;
							cell_logp += minus_log_sigma - 0.91893853320467274178032973640562 - d * d * inv_two_sigma_sq;
						} else if (pr.error_model == 1) {
							cell_logp += logpdf_tnu4(y, x, mv.stdev);
						} else {
							const double d = y - x;
							cell_logp += min_log_sigma[(size_t)l][(size_t)k * n_sim + j] - 0.91893853320467274178032973640562 -
							             d * d * inv_two_sigma_sq_cell[(size_t)l][(size_t)k * n_sim + j];
						}
					}
					if (cell_logp == ninf) break; // .cpp:485-488: negative infinity -- no need to go further
				}
				cell_likelihoods[(size_t)i * n + j] = cell_logp;
				if (cell_logp != cell_logp) { // .cpp:301-304
					*logp_out = ninf;
					return;
				}
				if (cell_logp > ninf) finite_count++;
			}
			if (finite_count < n_obs) { // .cpp:316-320
				*logp_out = ninf;
				return;
			}
		}
		std::vector<double> cost((size_t)n_obs * n_sim);
		for (int i = 0; i < n_obs; i++)
			for (int j = 0; j < n_sim; j++) cost[(size_t)i * n_sim + j] = -cell_likelihoods[(size_t)i * n + j];
		const std::vector<int> matching = hungarian_match(n, n_sim, n_obs, cost); // .cpp:323
		if ((int)matching.size() != n_obs) {
			*logp_out = ninf;
			return;
		}
		double logp = 0.0;
		for (int i = 0; i < n_obs; i++) {
			const int j = matching[i];
			if (j == -1) {
				*logp_out = ninf;
				return;
			}
			logp += cell_likelihoods[(size_t)i * n + j];
		}
		*logp_out = logp * pr.weight;
		return;
	}

	if (pr.data_kind == 2) {
		// <data type="time_points">: DataLikelihoodTimePoints::Evaluate (DataLikelihoodTimePoints.cpp:209-345) -- at every timepoint
		// its own set of observed cells (row i of `observed`, NaN = no such cell at that time), matched to the simulated cells that
		// have a value there (in marker 0). synchronize="none". The Hungarian call is rectangular (observed cells x simulated cells
		// with a value); the reference's implementation keeps only the edges whose right node index lies below the number of LEFT
		// nodes (hungarian.cpp:81), i.e. the first simulated cells -- the adapter hands it the edge list as Evaluate builds it.
		const int rel = pr.value_relative_to_timepoint_ix;
		const int n_sim_all = ncell; // cell_trajectories has one entry per cell object (GetMaxNumberOfCells)
		double logp = 0.0;
		for (int ti = 0; ti < T; ti++) {
			std::vector<int> rows, cols;
			for (int i = 0; i < R; i++) { // a row with any finite value in any marker, .cpp:222-227
				bool any = false;
				for (int l = 0; l < L; l++) any = any || std::isfinite(markers[(size_t)l].observed[(size_t)i * T + ti]);
				if (any) rows.push_back(i);
			}
			if (rows.empty()) continue;
			for (int j = 0; j < n_sim_all; j++) {
				if (pr.use_only_nondivided && j >= ninit) continue; // NotifySimulatedValue skips newborn cells: their entries stay NaN
				if (!std::isnan(xs[(size_t)ti * ncell + j]) && (rel < 0 || !std::isnan(xs[(size_t)rel * ncell + j]))) cols.push_back(j);
			}
			if (cols.size() < rows.size()) { // .cpp:241-245
				*logp_out = -std::numeric_limits<double>::infinity();
				return;
			}
			const int fd = (int)rows.size(), fs = (int)cols.size();
			std::vector<double> lik((size_t)fd * fs), cost((size_t)fd * fs);
			for (int a2 = 0; a2 < fd; a2++)
				for (int b2 = 0; b2 < fs; b2++) {
					double cell_logp = 0.0;
					for (int l = 0; l < L; l++) {
						const MarkerView& mv = markers[(size_t)l];
						double x = mv.xs[(size_t)ti * ncell + cols[b2]];
						if (rel >= 0) {
							x += mv.offset;
							x /= mv.xs[(size_t)rel * ncell + cols[b2]];
							x *= mv.scale;
						} else {
							x *= mv.scale;
							x += mv.offset;
						}
						const double y = mv.observed[(size_t)rows[a2] * T + ti];
						if (std::isnan(y)) continue; // .cpp:275-279: the other markers' missing values are ignored
						if (pr.error_model == 0) cell_logp += logpdf_normal(y, x, mv.stdev);
						else if (pr.error_model == 1) cell_logp += logpdf_tnu4(y, x, mv.stdev);
						else cell_logp = nan; // assert(false) in the reference: the other error models do not exist for this data type
					}
					lik[(size_t)a2 * fs + b2] = cell_logp;
					cost[(size_t)a2 * fs + b2] = -cell_logp;
				}
			const std::vector<int> matching = hungarian_match(fd, fs, fd, cost); // .cpp:306-307
			if ((int)matching.size() != fd) {
				*logp_out = -std::numeric_limits<double>::infinity();
				return;
			}
			for (int a2 = 0; a2 < fd; a2++) {
				if (matching[a2] == -1) {
					*logp_out = -std::numeric_limits<double>::infinity();
					return;
				}
				logp += lik[(size_t)a2 * fs + matching[a2]];
			}
		}
		*logp_out = logp * pr.weight;
		return;
	}

	// Evaluate (.cpp:85-159)
	double stdev = (pr.stdev_ix >= 0) ? transformed[pr.stdev_ix] : pr.stdev;
	const double offset = (pr.offset_ix >= 0) ? transformed[pr.offset_ix] : pr.offset;
	const double scale = (pr.scale_ix >= 0) ? transformed[pr.scale_ix] : pr.scale;
	if (pr.stdev_relative_to_scale) stdev *= scale; // GetCurrentSTDev, DataLikelihoodBase.cpp:151-153
	const double prop_stdev = (pr.proportional_stdev_ix >= 0) ? transformed[pr.proportional_stdev_ix] : pr.proportional_stdev;
	const double minus_log_sigma = -log(stdev), inv_two_sigma_sq = 1.0 / (2.0 * stdev * stdev);
	if (pr.relative_to_time_average) {
		// .cpp:105-112: offset first, then the logarithm of every value relative to the average over the timepoints, then scale
		double time_average = 0.0;
		for (int i = 0; i < T; i++) {
			population_average[i] += offset;
			time_average += population_average[i];
		}
		time_average /= (double)T;
		for (int i = 0; i < T; i++) {
			population_average[i] = log(population_average[i] / time_average);
			population_average[i] *= scale;
		}
	} else {
		for (int i = 0; i < T; i++) {
			population_average[i] *= scale;
			population_average[i] += offset;
		}
	}
	double logp = 0.0;
	for (int i = 0; i < T; i++) {
		const double x = population_average[i];
		if (std::isnan(x)) {
			double first_ok = pr.timepoints[T - 1], last_ok = pr.timepoints[0];
			for (int m = 0; m < T; m++) if (!std::isnan(population_average[m])) { first_ok = pr.timepoints[m]; break; }
			for (int m = T - 1; m >= 0; m--) if (!std::isnan(population_average[m])) { last_ok = pr.timepoints[m]; break; }
			const double time_offset = std::min(std::abs(pr.timepoints[i] - first_ok), std::abs(pr.timepoints[i] - last_ok));
			const double pen = (pr.error_model == 1) ? logpdf_tnu4(time_offset, 0, pr.missing_stdev) : logpdf_normal(time_offset, 0, pr.missing_stdev);
			for (int j = 0; j < R; j++) if (!std::isnan(pr.observed[(size_t)j * T + i])) logp += pen;
		} else {
			for (int j = 0; j < R; j++) {
				const double obs = pr.observed[(size_t)j * T + i];
				if (std::isnan(obs)) continue;
				// EvaluateValue(observed_data(j, i), x, 0): first argument is named `simulated` (argument swap)
				// so inside EvaluateValue (DataLikelihoodTimeCourseBase.cpp:267-299) `simulated` = obs and `observed` = x
				if (pr.error_model == 1) {
					logp += logpdf_tnu4(x, obs, stdev);
				} else if (pr.error_model == 2) {
					logp += logpdf_normal(x, obs, prop_stdev * std::max(obs, 0.0));
				} else if (pr.error_model == 3) {
					logp += logpdf_normal(x, obs, stdev + prop_stdev * std::max(obs, 0.0));
				} else {
					const double d = x - obs;
					logp += minus_log_sigma - 0.91893853320467274178032973640562 - d * d * inv_two_sigma_sq;
				}
			}
		}
	}
	*logp_out = logp * pr.weight;
}

template <class Solver>
int evaluate(const oracle_cellpop_problem* prob, size_t num_chains, const double* values, double* logp, double* cell_values,
             int32_t* cell_steps, double* population_average, int num_threads, int64_t* cell_counters = nullptr)
{
	if (!prob || !values || !logp || !prob->derivative) return -1;
	const size_t T = prob->num_timepoints, nc = prob->divide_cells ? prob->max_cells : prob->num_cells, nvar = prob->num_variables;
	if (num_threads < 1) num_threads = 1;
	// threads that the chains cannot use go to the cells of a chain (results do not depend on the split: every cell writes its
	// own column, the sums run in cell order afterwards)
	int cell_threads = 1;
	if ((size_t)num_threads > num_chains) {
		cell_threads = (int)((size_t)num_threads / (num_chains ? num_chains : 1));
		num_threads = (int)num_chains;
	}
	std::atomic<size_t> next(0);
	auto worker = [&]() {
		for (;;) {
			size_t c = next.fetch_add(1);
			if (c >= num_chains) break;
			evaluate_chain<Solver>(*prob, values + c * nvar, logp + c, cell_values ? cell_values + c * T * nc : nullptr,
			                       cell_steps ? cell_steps + c * nc : nullptr, population_average ? population_average + c * T : nullptr,
			                       cell_counters ? cell_counters + c * nc * ORACLE_NUM_COUNTERS : nullptr, cell_threads);
		}
	};
	if (num_threads == 1) {
		worker();
	} else {
		std::vector<std::thread> th;
		for (int i = 0; i < num_threads; i++) th.emplace_back(worker);
		for (auto& t : th) t.join();
	}
	return 0;
}

} // namespace cellpop_glue
