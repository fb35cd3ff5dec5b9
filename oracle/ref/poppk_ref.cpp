// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product.
//
// Thin restatement of the Boost/NetCDF-bound glue of
//   LikelihoodPopPKTrajectory::EvaluateLogProbability  (src/likelihoods/LikelihoodPopPKTrajectory.cpp:259-444)
// around the reference's REAL, unmodified, compiled solver stack:
//   ODESolverCVODE (src/odecommon/ODESolverCVODE.cpp) -> vendored CVODE 5.3.0
//   with the Eigen N_Vector / SUNMatrix / SUNLinearSolver (src/odecommon/*_eigen.cpp).
// Everything numerical below the per-patient loop is the reference itself; what
// is restated here (because the original translation unit needs Boost + NetCDF):
//   - parameter construction            cpp:263-330
//   - discontinuity (dosing) callbacks  cpp:644-690
//   - RHS / analytic Jacobians          cpp:446-494
//   - Student-t4 observation log-pdf    src/utils/ProbabilityDistributions.cpp:216-224
//   - VariableSet::TransformVariable    src/sampler/VariableSet.cpp:97-124
//   - bcm3::QuantileNormal              src/utils/ProbabilityDistributions.cpp:359-363
//     (boost::math::quantile(normal) = mu - sigma*sqrt(2)*erfc_inv(2p); Boost is
//      not vendored and not installed, so Phi^-1 is computed here by a rational
//      start + two Halley steps on libm erfc -- full double accuracy, checked
//      against scipy.special.ndtri in tests/test_oracle.py)
// The per-patient memo cache (cpp:332-353,429-436) is an optimisation that never
// changes results and is not reproduced.
#include "Utils.h"
#include "ODESolverCVODE.h"
#include <cvode/cvode.h>
#include <cvode/cvode_ls.h>
#include <atomic>

#include "../oracle_api.h"

namespace {

// subclass only to reach the protected cvode_mem for the statistics counters
class ProbedSolver : public ODESolverCVODE {
public:
	void AccumulateCounters(int64_t* cnt) const
	{
		long int v;
		if (CVodeGetNumRhsEvals(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NFE] += v;
		if (CVodeGetNumLinSolvSetups(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NSETUPS] += v;
		if (CVodeGetNumErrTestFails(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NETF] += v;
		if (CVodeGetNumNonlinSolvConvFails(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NCFN] += v;
		if (CVodeGetNumNonlinSolvIters(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NNI] += v;
		if (CVodeGetNumJacEvals(cvode_mem, &v) == 0) cnt[ORACLE_CNT_NJE] += v;
	}
};

double ndtri(double p)
{
	if (!(p > 0.0 && p < 1.0)) {
		if (p == 0.0) return -std::numeric_limits<double>::infinity();
		if (p == 1.0) return std::numeric_limits<double>::infinity();
		return std::numeric_limits<double>::quiet_NaN();
	}
	static const double a[6] = { -3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
	                             1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00 };
	static const double b[5] = { -5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
	                             6.680131188771972e+01, -1.328068155288572e+01 };
	static const double c[6] = { -7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
	                             -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00 };
	static const double d[4] = { 7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
	                             3.754408661907416e+00 };
	const double plow = 0.02425;
	double x;
	if (p < plow) {
		double q = sqrt(-2.0 * log(p));
		x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
		    ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
	} else if (p > 1.0 - plow) {
		double q = sqrt(-2.0 * log(1.0 - p));
		x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
		    ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
	} else {
		double q = p - 0.5;
		double r = q * q;
		x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
		    (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
	}
	for (int it = 0; it < 2; it++) {
		// Halley step on Phi(x) - p; the tail is taken on the side that avoids cancellation
		double e;
		if (x < 0.0) {
			e = 0.5 * erfc(-x * M_SQRT1_2) - p;
		} else {
			e = (1.0 - p) - 0.5 * erfc(x * M_SQRT1_2);
		}
		double u = e * 2.5066282746310002 * exp(0.5 * x * x);
		x = x - u / (1.0 + 0.5 * x * u);
	}
	return x;
}

inline double QuantileNormal(double p, double mu, double sigma)
{
	// boost::math::quantile(normal_distribution(mu, sigma), p): result = -erfc_inv(2p); result *= sigma*sqrt(2); result += mu
	double r = ndtri(p);
	r *= sigma;
	r += mu;
	return r;
}

inline double TransformVariable(int transform, double x)
{
	switch (transform) {
	case ORACLE_TRANSFORM_NONE:
		return x;
	case ORACLE_TRANSFORM_LOG:
		return exp(x);
	case ORACLE_TRANSFORM_LOG10:
		return bcm3::fastpow10(x);
	case ORACLE_TRANSFORM_LOGIT:
		if (x > 0) {
			double z = exp(-x);
			return 1.0 / (1.0 + z);
		} else {
			double z = exp(x);
			return z / (1.0 + z);
		}
	default:
		return x;
	}
}

inline double LogPdfTnu4(double x, double mu, double sigma)
{
	double xn = (x - mu) / sigma;
	return -0.9808292530117262 - 2.5 * log1p(0.25 * xn * xn) - log(sigma);
}

struct ParallelData {
	double dose, dosing_interval, dose_after_dose_change, dose_change_time;
	unsigned int intermittent;
	uint32_t skipped_days;
	double k_absorption, k_excretion, k_elimination, k_vod, k_periphery_fwd, k_periphery_bwd;
	double k_transit, n_transit, k_biphasic_switch_time, k_absorption2, last_treatment;
	bool biphasic_switch;
	double current_dose_time;
	ProbedSolver* solver;
	int64_t* counters; // may be null
};

inline bool CheckGiveTreatment(double t, const ParallelData& pd)
{
	bool give_treatment = true;
	int day = static_cast<int>(floor(t / 24.0));
	if (day >= 0 && day < 29 && ((pd.skipped_days >> day) & 1u)) {
		give_treatment = false;
	}
	if (pd.intermittent == 1) {
		double time_in_week = t - 7.0 * 24.0 * floor(t / (7.0 * 24.0));
		if (time_in_week >= 5.0 * 24.0) give_treatment = false;
	} else if (pd.intermittent == 2) {
		double time_in_treatment_course = t - 28.0 * 24.0 * floor(t / (28.0 * 24.0));
		if (time_in_treatment_course >= 21.0 * 24.0) give_treatment = false;
	} else if (pd.intermittent == 3) {
		double time_in_week = t - 7.0 * 24.0 * floor(t / (7.0 * 24.0));
		if (time_in_week >= 4.0 * 24.0) give_treatment = false;
	}
	return give_treatment;
}

void EvaluateChain(const oracle_poppk_problem& pr, const double* values, double* logp_out, double* conc,
                   double* patient_ll, int64_t* counters)
{
	const int P = pr.num_patients;
	const int T = pr.num_timepoints;
	const bool two = ORACLE_PK_IS_TWO(pr.pk_type);
	const bool biphasic = ORACLE_PK_IS_BIPHASIC(pr.pk_type), transit = ORACLE_PK_IS_TRANSIT(pr.pk_type);
	static const size_t pk_params_of_type[6] = { 4, 6, 7, 7, 6, 8 }; // cpp:99-120
	const size_t num_pk_params = pk_params_of_type[pr.pk_type];
	const size_t num_pk_pop_params = 2;
	const bool report_all = conc || patient_ll || counters;

	ParallelData pd;
	ProbedSolver solver;
	pd.solver = &solver;

	// CalculateDerivative_* / CalculateJacobian_* of all six model types (cpp:446-642): the biphasic types switch the
	// absorption rate, the transit types replace the bolus by a gamma-shaped input into the depot
	ODESolver::TDeriviativeFunction deriv = [&pd, two, biphasic, transit](OdeReal t, const OdeReal* y, OdeReal* dydt, void*) {
		double ka = pd.k_absorption;
		if (biphasic && !pd.biphasic_switch) ka = pd.k_absorption2;
		double input = 0.0;
		if (transit) {
			Real dose = pd.dose;
			if (t >= pd.dose_change_time) dose = pd.dose_after_dose_change;
			Real t_since_treatment = t - pd.last_treatment;
			Real log_n_transit_factorial = 0.9189385332046727 + (pd.n_transit + 0.5) * log(pd.n_transit) - pd.n_transit + log(1 + 1 / (12.0 * pd.n_transit));
			Real tr = exp((pd.n_transit * log(pd.k_transit * t_since_treatment) - pd.k_transit * t_since_treatment) - log_n_transit_factorial);
			input = pd.k_transit * tr * dose;
		}
		if (transit) dydt[0] = input - (ka + pd.k_excretion) * y[0];
		else dydt[0] = -(ka + pd.k_excretion) * y[0];
		if (!two) {
			dydt[1] = ka * y[0] - pd.k_elimination * y[1];
		} else {
			dydt[1] = ka * y[0] - pd.k_elimination * y[1] - pd.k_periphery_fwd * y[1] + pd.k_periphery_bwd * y[2];
			dydt[2] = pd.k_periphery_fwd * y[1] - pd.k_periphery_bwd * y[2];
		}
		return true;
	};
	ODESolver::TJacobianFunction jac = [&pd, two, biphasic](OdeReal t, const OdeReal* y, const OdeReal* dydt, OdeMatrixReal& jac, void*) {
		double ka = pd.k_absorption;
		if (biphasic && !pd.biphasic_switch) ka = pd.k_absorption2;
		jac(0, 0) = -(ka + pd.k_excretion);
		jac(1, 0) = ka;
		if (!two) {
			jac(1, 1) = -pd.k_elimination;
		} else {
			jac(1, 1) = -(pd.k_elimination + pd.k_periphery_fwd);
			jac(1, 2) = pd.k_periphery_bwd;
			jac(2, 1) = pd.k_periphery_fwd;
			jac(2, 2) = -pd.k_periphery_bwd;
		}
		return true;
	};
	solver.SetDerivativeFunction(deriv);
	solver.SetJacobianFunction(jac);
	solver.Initialize(two ? 3 : 2, NULL, 0);
	solver.SetSolverParameter("max_steps", pr.max_steps, 0.0);
	solver.SetTolerance(pr.rtol, pr.atol);

	ODESolver::TDiscontinuityCallback treatment_cb = [&pd, transit](OdeReal t, void*) -> Real {
		// LikelihoodPopPKTrajectory::TreatmentCallback, cpp:673-690 (called right before CVodeReInit)
		if (pd.counters) pd.solver->AccumulateCounters(pd.counters);
		pd.current_dose_time += pd.dosing_interval;
		if (CheckGiveTreatment(t, pd)) {
			double dose = pd.dose;
			if (t >= pd.dose_change_time) dose = pd.dose_after_dose_change;
			if (transit) pd.last_treatment = t;
			else pd.solver->set_current_y(0, pd.solver->get_current_y(0) + dose);
		}
		return pd.current_dose_time;
	};
	ODESolver::TDiscontinuityCallback treatment_cb_biphasic = [&pd](OdeReal t, void*) -> Real {
		// LikelihoodPopPKTrajectory::TreatmentCallbackBiphasic, cpp:692-718
		if (pd.counters) pd.solver->AccumulateCounters(pd.counters);
		if (pd.biphasic_switch) {
			pd.biphasic_switch = false;
			pd.current_dose_time += pd.dosing_interval;
			return pd.current_dose_time;
		} else {
			if (CheckGiveTreatment(t, pd)) {
				double dose = pd.dose;
				if (t >= pd.dose_change_time) dose = pd.dose_after_dose_change;
				pd.solver->set_current_y(0, pd.solver->get_current_y(0) + dose);
				pd.biphasic_switch = true;
				return pd.current_dose_time + pd.k_biphasic_switch_time;
			} else {
				pd.current_dose_time += pd.dosing_interval;
				return pd.current_dose_time;
			}
		}
	};

	Eigen::Map<const OdeVectorReal> time(pr.time, T);
	const double inf = std::numeric_limits<double>::infinity();

	double logp = 0.0;
	bool stopped = false;

	const size_t sdix = (size_t)pr.sd_ix;
	double sd = TransformVariable(pr.transforms[sdix], values[sdix]);
	double sd2 = TransformVariable(pr.transforms[sdix + 1], values[sdix + 1]);

	for (int j = 0; j < P; j++) {
		if (report_all) {
			if (conc)
				for (int i = 0; i < T; i++) conc[(size_t)j * T + i] = std::numeric_limits<double>::quiet_NaN();
			if (counters)
				for (int k = 0; k < ORACLE_NUM_COUNTERS; k++) counters[(size_t)j * ORACLE_NUM_COUNTERS + k] = 0;
		}
		if (stopped && !report_all) break;

		pd.dose = pr.dose[j];
		pd.dosing_interval = pr.dosing_interval[j];
		pd.dose_after_dose_change = pr.dose_after_dose_change[j];
		pd.dose_change_time = pr.dose_change_time[j];
		pd.intermittent = (unsigned int)pr.intermittent[j];
		if (pr.single) pd.intermittent = pr.intermittent[j] ? 1u : 0u;
		pd.skipped_days = pr.skipped_days[j];
		pd.counters = counters ? counters + (size_t)j * ORACLE_NUM_COUNTERS : nullptr;

		if (!pr.single) pd.k_absorption = bcm3::fastpow10(QuantileNormal(values[num_pk_params + num_pk_pop_params * (j + 1) + 0], values[0], values[num_pk_params + 0]));
		pd.k_excretion = TransformVariable(pr.transforms[1], values[1]);
		pd.k_vod = std::isnan(pr.fixed_vod) ? TransformVariable(pr.transforms[3], values[3]) : pr.fixed_vod;
		if (pr.single) {
			pd.k_absorption = TransformVariable(pr.transforms[0], values[0]);
			pd.k_elimination = TransformVariable(pr.transforms[2], values[2]) / pd.k_vod;
		} else
			pd.k_elimination = bcm3::fastpow10(QuantileNormal(values[num_pk_params + num_pk_pop_params * (j + 1) + 1], values[2], values[num_pk_params + 1])) / pd.k_vod;
		if (two) {
			if (std::isnan(pr.fixed_periphery_fwd)) {
				pd.k_periphery_fwd = TransformVariable(pr.transforms[4], values[4]);
				pd.k_periphery_bwd = TransformVariable(pr.transforms[5], values[5]);
			} else {
				pd.k_periphery_fwd = pr.fixed_periphery_fwd;
				pd.k_periphery_bwd = pr.fixed_periphery_bwd;
			}
		}

		if (transit) { // cpp:296-301
			pd.n_transit = TransformVariable(pr.transforms[pr.n_transit_ix], values[pr.n_transit_ix]);
			pd.k_transit = (pd.n_transit + 1) / TransformVariable(pr.transforms[pr.mean_transit_time_ix], values[pr.mean_transit_time_ix]);
		}
		if (biphasic) { // cpp:302-310
			pd.k_biphasic_switch_time = TransformVariable(pr.transforms[pr.biphasic_uptake_time_ix], values[pr.biphasic_uptake_time_ix]);
			if (!pr.single) pd.k_biphasic_switch_time = std::min(pd.k_biphasic_switch_time, pd.dosing_interval - 1e-2);
			pd.k_absorption2 = TransformVariable(pr.transforms[pr.mean_absorption2_ix], values[pr.mean_absorption2_ix]);
		}
		pd.last_treatment = 0.0;

		if (biphasic) { // cpp:357-364
			pd.biphasic_switch = true;
			pd.current_dose_time = 0;
			solver.SetDiscontinuity(pd.k_biphasic_switch_time, treatment_cb_biphasic, nullptr);
		} else {
			pd.current_dose_time = pd.dosing_interval;
			solver.SetDiscontinuity(pd.dosing_interval, treatment_cb, nullptr);
		}

		OdeVectorReal initial_conditions(two ? 3 : 2);
		initial_conditions.setZero();
		initial_conditions[0] = transit ? 0.0 : pd.dose; // cpp:366-373

		double conversion = (1e6 / pr.mol_weight) / pd.k_vod;

		OdeVectorReal simulate_time = time.segment(0, pr.simulate_until[j]);
		OdeMatrixReal simulated_trajectories;

		double patient_logllh = 0.0;
		if (simulate_time.size() > 0) {
			bool ok = solver.SolveReturnSolution(initial_conditions, &simulate_time, &simulated_trajectories);
			if (counters) {
				solver.AccumulateCounters(pd.counters);
				pd.counters[ORACLE_CNT_STEPS] = (int64_t)solver.GetNumSteps();
				pd.counters[ORACLE_CNT_OK] = ok ? 1 : 0;
			}
			if (!ok) {
				patient_logllh = -inf;
			} else {
				for (ptrdiff_t i = 0; i < simulate_time.size(); i++) {
					double x = conversion * simulated_trajectories(1, i);
					double y = pr.observed_concentration[(size_t)j * T + i];
					if (conc) conc[(size_t)j * T + i] = x;
					if (!std::isnan(y)) {
						patient_logllh += LogPdfTnu4(x, y, sd + sd2 * std::max(x, 0.0));
					}
					if (std::isnan(x) && !pr.single) {
						patient_logllh = -inf;
						break;
					}
				}
			}
		}
		if (patient_ll) patient_ll[j] = patient_logllh;

		if (!stopped) {
			logp += patient_logllh;
			if (logp == -inf) stopped = true;
		}
	}
	*logp_out = logp;
}

} // namespace

namespace refglue {
double ndtri(double p) { return ::ndtri(p); }
}

extern "C" int oracle_poppk_evaluate(const oracle_poppk_problem* prob, size_t num_chains, const double* values,
                                     double* logp, double* conc, double* patient_ll, int64_t* counters, int num_threads)
{
	if (!prob || !values || !logp) return -1;
	if (prob->pk_type < ORACLE_PK_ONE || prob->pk_type > ORACLE_PK_TWO_TRANSIT) return -2;
	const size_t P = (size_t)prob->num_patients, T = (size_t)prob->num_timepoints, nvar = (size_t)prob->num_variables;
	if (num_threads < 1) num_threads = 1;
	if ((size_t)num_threads > num_chains) num_threads = (int)num_chains;

	std::atomic<size_t> next(0);
	auto worker = [&]() {
		for (;;) {
			size_t c = next.fetch_add(1);
			if (c >= num_chains) break;
			EvaluateChain(*prob, values + c * nvar, logp + c, conc ? conc + c * P * T : nullptr,
			              patient_ll ? patient_ll + c * P : nullptr,
			              counters ? counters + c * P * ORACLE_NUM_COUNTERS : nullptr);
		}
	};
	if (num_threads == 1) {
		worker();
	} else {
		std::vector<std::thread> threads;
		for (int i = 0; i < num_threads; i++) threads.emplace_back(worker);
		for (auto& t : threads) t.join();
	}
	return 0;
}

extern "C" const char* oracle_kind(void) { return "ref"; }
