// GPU-backed drop-in for PharmacoLikelihoodPopulation (src/pharmaco/PharmacoLikelihoodPopulation.{h,cpp}) on top of the C ABI
// (include/bcm3b200.h, model kind "pharmaco_population"). likelihood.xml surface: <bcm_likelihood type="pharmaco_population">
// <pk_model drug= trial= peripheral_compartment= num_transit_compartments= bioavailability= likelihood_cache_size=/> as in the
// reference (cpp:47-62); variables are found by name in the prior (PostInitialize, cpp:102-188). The NetCDF reader is out of
// scope: the trial arrays of pkdata.nc are supplied with SetTrialData() before PostInitialize().
// Constructed with single_patient = true the class is the drop-in for PharmacoLikelihoodSingle (src/pharmaco/
// PharmacoLikelihoodSingle.{h,cpp}; type="pharmaco_single", model kind "pharmaco_single"): <pk_model ... patient= biphasic_absorption=
// metabolite=>, the variables "absorption", "clearance", "volume_of_distribution", ... are the patient's rates themselves.
#pragma once

#include "Likelihood.h"
#include "LikelihoodPopPKTrajectoryB200.h"

class PharmacoLikelihoodPopulationB200 : public bcm3::Likelihood {
public:
	typedef LikelihoodPopPKTrajectoryB200::TrialData TrialData; // the same NetCDF group (Patient::Load, PharmacoPatient.cpp:24-46)

	PharmacoLikelihoodPopulationB200(size_t sampling_threads, size_t evaluation_threads, bool single_patient = false);
	void SetPatientIDs(const std::vector<std::string>& ids) { patient_ids = ids; } // the "patients" dimension of the trial (PharmacoPatient.cpp:14-18)
	void SetPatientID(const std::string& patient) { patient_id = patient; }        // the pharmacosingle.patient option
	~PharmacoLikelihoodPopulationB200() override;

	bool Initialize(std::shared_ptr<const bcm3::VariableSet> varset, const bcm3::XmlNode& likelihood_node) override;
	void SetTrialData(const TrialData& data) { trial = data; }
	void SetDevice(int dev) { device = dev; }
	bool PostInitialize() override;
	bool IsReentrant() override { return true; }
	bool EvaluateLogProbability(size_t threadix, const bcm3::VectorReal& values, bcm3::Real& logp) override;
	bool EvaluateLogProbabilityBatch(const bcm3::MatrixReal& values, bcm3::VectorReal& logp) override;
	const std::string& LastError() const { return last_error; }

private:
	std::shared_ptr<const bcm3::VariableSet> varset;
	std::string drug, trial_name, patient_id;
	std::vector<std::string> patient_ids;
	bool single = false, use_biphasic = false, use_metabolite = false;
	bool use_peripheral = false, use_bioavailability = false;
	size_t num_transit = 0;
	TrialData trial;
	void* handle = nullptr;
	int device = 0;
	std::vector<int> status;
	std::string last_error;
};
