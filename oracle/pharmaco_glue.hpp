// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Restatement of the Boost/NetCDF-bound glue of the pharmaco_population likelihood around a compartment-model adapter:
//   Patient::Load                                        src/pharmaco/PharmacoPatient.cpp:8-116 (dose schedule, observations with a value)
//   PharmacoLikelihoodPopulation::EvaluateLogProbability src/pharmaco/PharmacoLikelihoodPopulation.cpp:202-247
//   PharmacoLikelihoodPopulation::SetupSimulation        :259-340
// (the per-patient result cache of :356-393 returns what a recomputation would give and is left out).
// Used twice: oracle/ref/pharmaco_ref.cpp plugs in the reference's own PharmacokineticModel (compiled from its source with
// Eigen's matrix exponential), oracle/pharmaco_port.cpp a plain restatement. The adapter provides
//   void configure(bool peripheral, int num_transit);
//   void configure_single(bool biphasic_absorption, bool metabolite); void set_single(direct_absorption_rate, metabolite_conversion_rate);
//     -- PharmacoLikelihoodSingle (src/pharmaco/PharmacoLikelihoodSingle.cpp), the one-patient form of the same likelihood
//   bool solve(absorption, excretion, elimination, kf, kb, transit_rate, bioavailability, treat_times, treat_doses, obs_times, out)
#pragma once

#include <atomic>
#include <cmath>
#include <limits>
#include <thread>
#include <vector>

#include "oracle_api.h"

namespace pharmaco_glue {

double ndtri(double p); // defined by the including translation unit

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
inline double fastpow10(double x) { return exp(x * 2.3025850929940459); }
inline double logpdf_tnu4(double x, double mu, double sigma)
{
	double xn = (x - mu) / sigma;
	return -0.9808292530117262 - 2.5 * log1p(0.25 * xn * xn) - log(sigma);
}

struct PatientData { // Patient::Load
	std::vector<double> treatment_timepoints, treatment_doses, observation_timepoints, observed_concentrations;
	std::vector<int> grid; // index of every kept observation in the trial's time grid
};

inline PatientData load_patient(const oracle_pharmaco_problem& pr, int j)
{
	PatientData p;
	const int T = pr.num_timepoints;
	const double dose = pr.dose[j], dosing_interval = pr.dosing_interval[j], dac = pr.dose_after_dose_change[j], dct = pr.dose_change_time[j];
	const int intermittent = pr.intermittent[j];
	const double last_time = 696;
	double t = 0;
	while (t < last_time) {
		bool give_treatment = true;
		const int day = (int)floor(t / 24.0);
		if (day >= 0 && day < 29 && (pr.skipped_days[j] >> day) & 1u) give_treatment = false;
		if (intermittent == 1) {
			if (t - 7.0 * 24.0 * floor(t / (7.0 * 24.0)) >= 5.0 * 24.0) give_treatment = false;
		} else if (intermittent == 2) {
			if (t - 28.0 * 24.0 * floor(t / (28.0 * 24.0)) >= 21.0 * 24.0) give_treatment = false;
		} else if (intermittent == 3) {
			if (t - 7.0 * 24.0 * floor(t / (7.0 * 24.0)) >= 4.0 * 24.0) give_treatment = false;
		}
		if (give_treatment) {
			p.treatment_timepoints.push_back(t);
			p.treatment_doses.push_back((!std::isnan(dct) && t >= dct) ? dac : dose);
		}
		t += dosing_interval;
	}
	for (int i = 0; i < T; i++) {
		const double y = pr.observed_concentration[(size_t)j * T + i];
		if (!std::isnan(y)) {
			p.observation_timepoints.push_back(pr.time[i]);
			p.observed_concentrations.push_back(y);
			p.grid.push_back(i);
		}
	}
	return p;
}

template <class Model>
void evaluate_chain(const oracle_pharmaco_problem& pr, const std::vector<PatientData>& patients, const double* v, double* logp_out, double* conc,
                    double* patient_ll)
{
	const int P = pr.num_patients, T = pr.num_timepoints;
	const double nan = std::numeric_limits<double>::quiet_NaN();
	auto tv = [&](int ix) { return transform_variable(pr.transforms[ix], v[ix]); };
	auto marginal = [&](int mean_ix, int sigma_ix, const int32_t* pix, int j) {
		if (sigma_ix < 0) return fastpow10(v[mean_ix]);
		return fastpow10(v[mean_ix] + v[sigma_ix] * ndtri(v[pix[j]])); // QuantileNormal(p, mu, sigma)
	};
	Model model;
	model.configure(pr.use_peripheral != 0, pr.num_transit);
	model.configure_single(pr.single && pr.use_biphasic, pr.single && pr.use_metabolite); // PharmacoLikelihoodSingle.cpp:127-149
	double logp = 0.0;
	const double additive_sd = pr.additive_sd_ix >= 0 ? tv(pr.additive_sd_ix) : 0.0;
	const double proportional_sd = pr.proportional_sd_ix >= 0 ? tv(pr.proportional_sd_ix) : 0.0;
	std::vector<double> sim;
	for (int j = 0; j < P; j++) {
		const PatientData& pt = patients[j];
		// SetupSimulation
		// PharmacoLikelihoodSingle.cpp:163-178: the single-patient likelihood takes the transformed variables as the rates
		const double absorption = pr.single ? tv(pr.mean_absorption_ix) : marginal(pr.mean_absorption_ix, pr.sigma_absorption_ix, pr.p_absorption_ix, j);
		const double excretion = pr.mean_excretion_ix >= 0 ? (pr.single ? tv(pr.mean_excretion_ix) : marginal(pr.mean_excretion_ix, pr.sigma_excretion_ix, pr.p_excretion_ix, j)) : 0.0;
		const double clearance = pr.single ? tv(pr.mean_clearance_ix) : marginal(pr.mean_clearance_ix, pr.sigma_clearance_ix, pr.p_clearance_ix, j);
		const double vod = pr.single ? tv(pr.mean_vod_ix) : marginal(pr.mean_vod_ix, pr.sigma_vod_ix, pr.p_vod_ix, j);
		if (pr.single) model.set_single(pr.use_biphasic ? tv(pr.direct_absorption_ix) : 0.0, pr.use_metabolite ? tv(pr.metabolite_conversion_ix) : 0.0); // :192-199
		double kf = nan, kb = nan, transit_rate = nan, bioavailability = 1.0;
		if (pr.use_peripheral) {
			kf = tv(pr.periph_fwd_ix);
			kb = tv(pr.periph_bwd_ix);
		}
		if (pr.num_transit > 0) {
			const double transit_time = (pr.sigma_transit_ix < 0) ? tv(pr.mean_transit_time_ix)
			                                                       : fastpow10(v[pr.mean_transit_time_ix] + v[pr.sigma_transit_ix] * ndtri(v[pr.p_transit_ix[j]]));
			transit_rate = (pr.num_transit + 1.0) / transit_time;
		}
		if (pr.use_bioavailability) bioavailability = v[pr.p_bioavailability_ix[j]];
		const double conversion = (1e6 / pr.mol_weight) / vod;
		sim.assign(pt.observation_timepoints.size(), nan);
		double this_logp = 0.0;
		bool ok = pt.observation_timepoints.empty() ||
		          model.solve(absorption, excretion, clearance / vod, kf, kb, transit_rate, bioavailability, pt.treatment_timepoints, pt.treatment_doses,
		                      pt.observation_timepoints, sim);
		if (conc)
			for (int i = 0; i < T; i++) conc[(size_t)j * T + i] = nan;
		if (ok) {
			for (size_t ti = 0; ti < pt.observation_timepoints.size(); ti++) {
				const double x = conversion * sim[ti];
				if (conc) conc[(size_t)j * T + pt.grid[ti]] = x;
				if (std::isnan(x) || std::isinf(x)) {
					this_logp = -std::numeric_limits<double>::infinity();
					break;
				}
				const double y = pt.observed_concentrations[ti];
				if (!std::isnan(y)) this_logp += logpdf_tnu4(x, y, additive_sd + proportional_sd * std::max(x, 0.0));
			}
		} else {
			this_logp = -std::numeric_limits<double>::infinity();
		}
		if (patient_ll) patient_ll[j] = this_logp;
		logp += this_logp;
	}
	*logp_out = logp;
}

template <class Model>
int evaluate(const oracle_pharmaco_problem* prob, size_t num_chains, const double* values, double* logp, double* conc, double* patient_ll, int num_threads)
{
	if (!prob || !values || !logp) return -1;
	const size_t P = prob->num_patients, T = prob->num_timepoints, nvar = prob->num_variables;
	std::vector<PatientData> patients;
	for (size_t j = 0; j < P; j++) patients.push_back(load_patient(*prob, (int)j));
	if (num_threads < 1) num_threads = 1;
	if ((size_t)num_threads > num_chains) num_threads = (int)num_chains;
	std::atomic<size_t> next(0);
	auto worker = [&]() {
		for (;;) {
			const size_t c = next.fetch_add(1);
			if (c >= num_chains) break;
			evaluate_chain<Model>(*prob, patients, values + c * nvar, logp + c, conc ? conc + c * P * T : nullptr, patient_ll ? patient_ll + c * P : nullptr);
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

} // namespace pharmaco_glue
