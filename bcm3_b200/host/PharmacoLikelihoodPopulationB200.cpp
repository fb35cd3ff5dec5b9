#include "PharmacoLikelihoodPopulationB200.h"

extern "C" {
#include "bcm3b200.h"
}

using bcm3::Real;

PharmacoLikelihoodPopulationB200::PharmacoLikelihoodPopulationB200(size_t, size_t, bool single_patient) : single(single_patient) {}

PharmacoLikelihoodPopulationB200::~PharmacoLikelihoodPopulationB200()
{
	if (handle) bcm3b200_destroy(handle);
}

bool PharmacoLikelihoodPopulationB200::Initialize(std::shared_ptr<const bcm3::VariableSet> vs, const bcm3::XmlNode& node)
{
	varset = vs;
	const bcm3::XmlNode* model = node.child("pk_model");
	if (!model || !model->has("drug") || !model->has("trial")) { // cpp:49-52: both are required attributes
		last_error = "Error parsing likelihood file: pk_model needs drug and trial";
		return false;
	}
	drug = model->get("drug");
	trial_name = model->get("trial");
	use_peripheral = model->get_bool("peripheral_compartment", false);
	num_transit = (size_t)model->get_int("num_transit_compartments", 0);
	use_bioavailability = model->get_bool("bioavailability", false);
	if (single) { // PharmacoLikelihoodSingle::Initialize, PharmacoLikelihoodSingle.cpp:39-50
		use_bioavailability = false;
		patient_id = model->get("patient");
		use_biphasic = model->get_bool("biphasic_absorption", false);
		use_metabolite = model->get_bool("metabolite", false);
	}
	// likelihood_cache_size (cpp:55): the per-patient memo of previous results returns what a recomputation gives; not needed here
	return true;
}

bool PharmacoLikelihoodPopulationB200::PostInitialize()
{
	if (single) {
		// PharmacoLikelihoodSingle.cpp:57-60 and Patient::Load (PharmacoPatient.cpp:14-22): the patient has to be named and in the trial
		if (patient_id.empty()) {
			last_error = "Patient ID has not been specified in either the likelihood or as command-line option.";
			return false;
		}
		size_t ix = trial.dose.size();
		if (!patient_ids.empty()) {
			for (size_t i = 0; i < patient_ids.size(); i++)
				if (patient_ids[i] == patient_id) ix = i;
		} else {
			char* end = nullptr;
			const unsigned long long v = strtoull(patient_id.c_str(), &end, 10);
			if (end && *end == 0 && end != patient_id.c_str()) ix = (size_t)v;
		}
		if (ix >= trial.dose.size()) {
			last_error = "Cannot find patient \"" + patient_id + "\" in data file";
			return false;
		}
		const size_t Tn = trial.time.size();
		TrialData one;
		one.time = trial.time;
		one.observed_concentration.assign(trial.observed_concentration.begin() + ix * Tn, trial.observed_concentration.begin() + (ix + 1) * Tn);
		one.dose.assign(1, trial.dose[ix]);
		one.dosing_interval.assign(1, trial.dosing_interval[ix]);
		one.dose_after_dose_change.assign(1, trial.dose_after_dose_change[ix]);
		one.dose_change_time.assign(1, trial.dose_change_time[ix]);
		one.intermittent.assign(1, trial.intermittent[ix]);
		one.treatment_interruptions.assign(trial.treatment_interruptions.begin() + ix * 29, trial.treatment_interruptions.begin() + (ix + 1) * 29);
		trial = one;
	}
	const size_t P = trial.dose.size(), T = trial.time.size(), nvar = varset->GetNumVariables();
	const size_t none = std::numeric_limits<size_t>::max();
	std::string desc = "drug=" + drug + ";num_patients=" + std::to_string(P) + ";num_timepoints=" + std::to_string(T) + ";num_variables=" + std::to_string(nvar) +
	                   ";peripheral_compartment=" + (use_peripheral ? "1" : "0") + ";num_transit_compartments=" + std::to_string(num_transit) +
	                   ";bioavailability=" + (use_bioavailability ? "1" : "0") + ";device=" + std::to_string(device);
	// PostInitialize, cpp:102-188: every variable is looked up by name; which ones exist decides the model
	static const char* roles[][2] = { { "additive_sd", "additive_error_standard_deviation" }, { "proportional_sd", "proportional_error_standard_deviation" },
		                              { "mean_absorption", "mean_absorption" }, { "mean_excretion", "mean_excretion" }, { "mean_clearance", "mean_clearance" },
		                              { "mean_volume_of_distribution", "mean_volume_of_distribution" }, { "sigma_absorption", "sigma_absorption" },
		                              { "sigma_excretion", "sigma_excretion" }, { "sigma_clearance", "sigma_clearance" },
		                              { "sigma_volume_of_distribution", "sigma_volume_of_distribution" }, { "sigma_transit_time", "sigma_transit_time" },
		                              { "peripheral_forward_rate", "peripheral_forward_rate" }, { "peripheral_backward_rate", "peripheral_backward_rate" },
		                              { "mean_transit_time", "mean_transit_time" } };
	// PharmacoLikelihoodSingle::PostInitialize, PharmacoLikelihoodSingle.cpp:75-146
	static const char* single_roles[][2] = { { "additive_sd", "additive_error_standard_deviation" }, { "proportional_sd", "proportional_error_standard_deviation" },
		                                     { "absorption", "absorption" }, { "excretion", "excretion" }, { "clearance", "clearance" },
		                                     { "volume_of_distribution", "volume_of_distribution" }, { "peripheral_forward_rate", "peripheral_forward_rate" },
		                                     { "peripheral_backward_rate", "peripheral_backward_rate" }, { "mean_transit_time", "mean_transit_time" },
		                                     { "direct_absorption", "direct_absorption" }, { "metabolite_conversion_rate", "metabolite_conversion_rate" } };
	if (single) {
		for (const auto& r : single_roles) {
			const size_t ix = varset->GetVariableIndex(r[1]);
			if (ix != none) desc += std::string(";") + r[0] + "_ix=" + std::to_string(ix);
		}
		desc += std::string(";biphasic_absorption=") + (use_biphasic ? "1" : "0") + ";metabolite=" + (use_metabolite ? "1" : "0");
	} else {
		for (const auto& r : roles) {
			const size_t ix = varset->GetVariableIndex(r[1]);
			if (ix != none) desc += std::string(";") + r[0] + "_ix=" + std::to_string(ix);
		}
	}
	if (bcm3b200_create(single ? "pharmaco_single" : "pharmaco_population", desc.data(), desc.size(), 1, &handle) != BCM3B200_OK) {
		last_error = bcm3b200_last_error();
		return false;
	}
	std::vector<double> transforms(nvar);
	for (size_t i = 0; i < nvar; i++) transforms[i] = (double)varset->GetTransform(i);
	auto set = [&](const char* name, const std::vector<double>& v, std::vector<size_t> shape) {
		if (bcm3b200_set_data(handle, name, v.data(), shape.data(), (int)shape.size()) != BCM3B200_OK) {
			last_error = bcm3b200_last_error();
			return false;
		}
		return true;
	};
	bool ok = set("time", trial.time, { T }) && set("observed_concentration", trial.observed_concentration, { P, T }) && set("dose", trial.dose, { P }) &&
	          set("dosing_interval", trial.dosing_interval, { P }) && set("dose_after_dose_change", trial.dose_after_dose_change, { P }) &&
	          set("dose_change_time", trial.dose_change_time, { P }) && set("intermittent", trial.intermittent, { P }) &&
	          set("treatment_interruptions", trial.treatment_interruptions, { P, 29 }) && set("transforms", transforms, { nvar });
	if (!ok) return false;
	// InitializePatientMarginals, cpp:342-354: p<i>_<name> for every marginal whose sigma is in the prior (bioavailability: when enabled)
	auto marginal = [&](const char* name, const char* array, bool needed) {
		if (!needed) return true;
		std::vector<double> ixs(P);
		for (size_t i = 0; i < P; i++) {
			const std::string varname = "p" + std::to_string(i + 1) + "_" + name;
			const size_t ix = varset->GetVariableIndex(varname);
			if (ix == none) {
				last_error = std::string("Standard deviation found for \"") + name + "\", but could not find prior variable for \"" + varname + "\"";
				return false;
			}
			ixs[i] = (double)ix;
		}
		return set(array, ixs, { P });
	};
	auto has = [&](const char* v) { return !single && varset->GetVariableIndex(v) != none; };
	ok = marginal("absorption", "patient_absorption_ix", has("sigma_absorption")) &&
	     marginal("excretion", "patient_excretion_ix", has("sigma_excretion") && has("mean_excretion")) &&
	     marginal("clearance", "patient_clearance_ix", has("sigma_clearance")) &&
	     marginal("volume_of_distribution", "patient_volume_of_distribution_ix", has("sigma_volume_of_distribution")) &&
	     marginal("transit_time", "patient_transit_time_ix", has("sigma_transit_time") && num_transit > 0) &&
	     marginal("bioavailability", "patient_bioavailability_ix", use_bioavailability && !single);
	if (!ok) return false;
	if (bcm3b200_finalize(handle) != BCM3B200_OK) {
		last_error = bcm3b200_last_error();
		return false;
	}
	return true;
}

bool PharmacoLikelihoodPopulationB200::EvaluateLogProbability(size_t, const bcm3::VectorReal& values, Real& logp)
{
	int st = 0;
	if (bcm3b200_evaluate_batch(handle, 1, values.size(), values.data(), &logp, &st) != BCM3B200_OK) {
		last_error = bcm3b200_last_error();
		return false;
	}
	return true;
}

bool PharmacoLikelihoodPopulationB200::EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp)
{
	logp.assign(values.cols(), -bcm3::kInf);
	status.assign(values.cols(), 0);
	if (values.cols() == 0) return true;
	if (bcm3b200_evaluate_batch(handle, values.cols(), values.rows(), values.data.data(), logp.data(), status.data()) != BCM3B200_OK) {
		last_error = bcm3b200_last_error();
		return false;
	}
	return true;
}
