#include "LikelihoodPopPKTrajectoryB200.h"

extern "C" {
#include "bcm3b200.h"
}

using bcm3::Real;

LikelihoodPopPKTrajectoryB200::LikelihoodPopPKTrajectoryB200(size_t, size_t, bool single_patient) : single(single_patient) {}

LikelihoodPopPKTrajectoryB200::~LikelihoodPopPKTrajectoryB200()
{
	if (handle) bcm3b200_destroy(handle);
}

bool LikelihoodPopPKTrajectoryB200::Initialize(std::shared_ptr<const bcm3::VariableSet> vs, const bcm3::XmlNode& node)
{
	varset = vs;
	const bcm3::XmlNode* model = node.child("pk_model");
	if (!model) {
		last_error = "Error parsing likelihood file: no pk_model";
		return false;
	}
	if (!model->has("drug") || !model->has("type")) {
		last_error = "Error parsing likelihood file: pk_model needs drug and type";
		return false;
	}
	drug = model->get("drug");
	pk_type_str = model->get("type");
	trial_name = model->get("trial");
	pkdata_file = model->get("pkdata_file");
	if (single && model->has("patient")) patient_id = model->get("patient"); // LikelihoodPharmacokineticTrajectory.cpp:96
	// cpp:64-67: parameters fixed in likelihood.xml instead of sampled; handed through as they are
	fixed_attributes.clear();
	for (const char* key : { "volume_of_distribution", "k_periphery_fwd", "k_periphery_bwd" })
		if (model->has(key)) fixed_attributes += std::string(";") + key + "=" + model->get(key);
	return true;
}

bool LikelihoodPopPKTrajectoryB200::PostInitialize()
{
	if (single) {
		// LikelihoodPharmacokineticTrajectory.cpp:111-114, 161-166: the patient has to be named and has to be in the trial
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
	const size_t sdix = varset->GetVariableIndex("standard_deviation"); // cpp:263
	if (sdix == std::numeric_limits<size_t>::max()) {
		last_error = "Could not find variable \"standard_deviation\"";
		return false;
	}
	std::string desc = "type=" + pk_type_str + ";drug=" + drug + ";num_patients=" + std::to_string(P) + ";num_timepoints=" +
	                   std::to_string(T) + ";num_variables=" + std::to_string(nvar) + ";sd_ix=" + std::to_string(sdix) +
	                   ";device=" + std::to_string(device0) + fixed_attributes;
	// the variables the variants look up by name (cpp:296-310)
	auto named = [&](const char* variable, const char* key) {
		const size_t ix = varset->GetVariableIndex(variable);
		if (ix == std::numeric_limits<size_t>::max()) {
			last_error = std::string("Could not find variable \"") + variable + "\"";
			return false;
		}
		desc += std::string(";") + key + "=" + std::to_string(ix);
		return true;
	};
	if (pk_type_str == "one_transit" || pk_type_str == "two_transit") {
		if (!named("n_transit", "n_transit_ix") || !named("mean_transit_time", "mean_transit_time_ix")) return false;
	} else if (!single && (pk_type_str == "one_biphasic_uptake" || pk_type_str == "two_biphasic_uptake")) { // positional (6, 7) in the single-patient likelihood
		if (!named("biphasic_uptake_time", "biphasic_uptake_time_ix") || !named("mean_absorption2", "mean_absorption2_ix")) return false;
	}
	if (bcm3b200_create(single ? "pharmacokinetic_trajectory" : "pop_pk_trajectory", desc.data(), desc.size(), num_devices, &handle) != BCM3B200_OK) {
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
	bool ok = set("time", trial.time, { T }) && set("observed_concentration", trial.observed_concentration, { P, T }) &&
	          set("dose", trial.dose, { P }) && set("dosing_interval", trial.dosing_interval, { P }) &&
	          set("dose_after_dose_change", trial.dose_after_dose_change, { P }) && set("dose_change_time", trial.dose_change_time, { P }) &&
	          set("intermittent", trial.intermittent, { P }) && set("treatment_interruptions", trial.treatment_interruptions, { P, 29 }) &&
	          set("transforms", transforms, { nvar });
	if (!ok) return false;
	if (bcm3b200_finalize(handle) != BCM3B200_OK) {
		last_error = bcm3b200_last_error();
		return false;
	}
	return true;
}

bool LikelihoodPopPKTrajectoryB200::EvaluateLogProbability(size_t, const bcm3::VectorReal& values, Real& logp)
{
	int st = 0;
	if (bcm3b200_evaluate_batch(handle, 1, values.size(), values.data(), &logp, &st) != BCM3B200_OK) {
		last_error = bcm3b200_last_error();
		return false;
	}
	return true; // a NaN stays in logp: the sampler turns it into an error (Sampler.cpp:172-178)
}

bool LikelihoodPopPKTrajectoryB200::EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp)
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
